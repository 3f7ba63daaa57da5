// CudaCanvas — the skity::Canvas subclass of the B200 backend.
//
// It overrides the same protected virtuals SWCanvas overrides
// (src/render/sw/sw_canvas.hpp:69-104) and keeps the same per-Save state, but
// instead of rasterising on the CPU it ENCODES each call into the flat display
// list of include/skb_dl.h; GPUSurface::Flush hands that list to the CUDA side
// through the C ABI (include/skb.h).  Host work is what both existing backends
// also do on the host: CTM / clip-bounds bookkeeping (inherited from
// skity::Canvas), stroke outline generation with the reference's own Stroke,
// temp-surface sizing for mask-filter blur, brush matrix composition.
#ifndef SKITY_B200_HOST_CUDA_CANVAS_HPP
#define SKITY_B200_HOST_CUDA_CANVAS_HPP

#include <functional>
#include <memory>
#include <skity/render/canvas.hpp>
#include <string>
#include <vector>

#include "skity_b200/host/dl_builder.hpp"

namespace skity {

class CudaCanvas : public Canvas {
 public:
  // `builder` outlives the canvas; `surface` is the display-list surface drawn into.
  CudaCanvas(skb::DlBuilder* builder, uint32_t surface, uint32_t width, uint32_t height);
  ~CudaCanvas() override = default;

  // First unsupported feature met while encoding ("" if none).  Unsupported
  // draws are dropped, never approximated on the CPU.
  const std::string& Unsupported() const { return unsupported_; }

  // The reference's sub-canvas plumbing (SWCanvas, sw_canvas.hpp:106-182): a layer canvas shares the
  // root's CTM stack and global clip bounds, and sees them through its own device-space offset.
  CanvasState* GetCanvasState() const override {
    return parent_canvas_ ? parent_canvas_->GetCanvasState() : Canvas::GetCanvasState();
  }
  const Rect& GetGlobalClipBounds() const override {
    return parent_canvas_ ? parent_canvas_->GetGlobalClipBounds() : Canvas::GetGlobalClipBounds();
  }

 protected:
  void OnClipRect(const Rect& rect, ClipOp op) override;
  void OnClipPath(const Path& path, ClipOp op) override;
  void OnDrawPath(const Path& path, const Paint& paint) override;
  void OnDrawPaint(const Paint& paint) override;
  void OnSaveLayer(const Rect& bounds, const Paint& paint) override;
  void OnDrawBlob(const TextBlob* blob, float x, float y, Paint const& paint) override;
  void OnDrawGlyphs(uint32_t count, const GlyphID* glyphs, const float* position_x,
                    const float* position_y, const Font& font, const Paint& paint) override;
  void OnDrawImageRect(std::shared_ptr<Image> image, const Rect& src, const Rect& dst,
                       const SamplingOptions& sampling, Paint const* paint) override;
  void OnSave() override;
  void OnRestore() override;
  void OnRestoreToCount(int saveCount) override;
  void OnFlush() override;
  uint32_t OnGetWidth() const override { return width_; }
  uint32_t OnGetHeight() const override { return height_; }

 private:
  struct State {
    uint32_t clip_id = 0;  // 0 = no clip spans (SWCanvas::State::HasClip() == false)
    bool has_layer = false;
  };
  // SWCanvas::LayerState (sw_canvas.hpp:48-63): an offscreen surface + the canvas that draws into it
  struct LayerState {
    Rect rel_bounds = {};
    Rect log_bounds = {};
    uint32_t surface = 0, width = 0, height = 0;
    std::unique_ptr<CudaCanvas> canvas;
    Paint paint = {};
  };
  static constexpr uint32_t kNoSurface = 0xFFFFFFFFu;  // a zero-sized layer: draws into it vanish

  Matrix CurrentTransform() const;
  Rect ScanClipBounds() const;
  void FillPath(const Path& path, const Paint& paint, bool stroke);
  void EmitFill(const Path& path, const Matrix& ctm, uint32_t paint_index);
  void EmitFillOp(uint32_t path_index, Path::PathFillType fill_type, const Matrix& ctm, uint32_t paint_index);
  uint32_t MakeBrush(const Paint& paint, bool stroke);
  uint32_t EncodeColorFilter(const Paint& paint);
  void HandleFilter(const Path& path, const Paint& paint);
  void HandleFilterOf(const Rect& source_bounds, const Paint& paint,
                      const std::function<void(CudaCanvas&, const Paint&)>& draw_source);
  void DrawSurfaceImage(uint32_t src_surface, uint32_t w, uint32_t h, const Rect& dst,
                        const Paint& paint, bool unpremul, const Matrix* shader_local = nullptr);
  LayerState* PeekLayerStack() { return layer_stack_.empty() ? nullptr : layer_stack_.back().get(); }
  void OnLayerRestore();
  bool IsDrawingLayer() const { return parent_canvas_ ? parent_canvas_->drawing_layer_ : drawing_layer_; }
  void SetDrawingLayer(bool v) { (parent_canvas_ ? parent_canvas_->drawing_layer_ : drawing_layer_) = v; }
  void NoteUnsupported(const char* what);

  skb::DlBuilder* builder_;
  uint32_t surface_;
  uint32_t width_;
  uint32_t height_;
  std::vector<State> state_stack_;
  std::vector<std::unique_ptr<LayerState>> layer_stack_;
  CudaCanvas* parent_canvas_ = nullptr;
  Vec2 global_offset_ = Vec2{0.f, 0.f};
  bool drawing_layer_ = false;
  std::string unsupported_;
};

// Lowers a Path into display-list segments exactly as Stroke::QuadPath followed
// by PathEdgeIter would traverse it (curves kept; they are flattened on the GPU).
void LowerPathToSegs(const Path& src, std::vector<skb_dl_seg>* out);

}  // namespace skity

#endif  // SKITY_B200_HOST_CUDA_CANVAS_HPP
