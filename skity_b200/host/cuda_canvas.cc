#include "skity_b200/host/cuda_canvas.hpp"

#include "src/effect/color_filter_base.hpp"
#include "src/effect/image_filter_base.hpp"
#include "src/effect/pixmap_shader.hpp"
#include "src/graphic/color_priv.hpp"
#include "src/tracing.hpp"  // SKITY_TRACE_EVENT: the reference's own hooks (sw_canvas.cc:298,316,358,414,442,644), same build switch

#include <cmath>
#include <functional>
#include <cstring>
#include <skity/effect/image_filter.hpp>
#include <skity/effect/mask_filter.hpp>
#include <skity/effect/path_effect.hpp>
#include <skity/effect/shader.hpp>
#include <skity/geometry/stroke.hpp>
#include <skity/graphic/bitmap.hpp>
#include <skity/graphic/image.hpp>
#include <skity/graphic/paint.hpp>
#include <skity/graphic/path.hpp>

namespace skity {

// ---------------------------------------------------------------------------
// Path lowering
// ---------------------------------------------------------------------------
namespace {

struct LPoint {
  float x, y;
  bool cubic_end;  // value is the source p3; the real point is the computed end of that cubic
};

// The destination path Stroke::QuadPath builds (src/geometry/stroke.cc:914-962), with
// conics and cubics kept as curves.  MoveTo / Close / InjectMoveToIfNeed follow
// src/graphic/path.cc:494-506,837-860,1327-1339.
struct LoweredPath {
  struct Item {
    uint8_t verb;   // Path::Verb numbering
    float p[8];     // curve-maths points handed out by Path::Iter (p0 = source previous point)
    float w;
  };
  std::vector<Item> items;
  std::vector<LPoint> points;
  std::vector<uint8_t> verbs;
  int32_t last_move_to_index = ~0;

  void MoveTo(float x, float y) {
    if (!verbs.empty() && verbs.back() == 0) {
      points.back() = LPoint{x, y, false};
      items.back().p[0] = x;
      items.back().p[1] = y;
    } else {
      last_move_to_index = static_cast<int32_t>(points.size());
      verbs.push_back(0);
      points.push_back(LPoint{x, y, false});
      Item it{};
      it.verb = 0;
      it.p[0] = x;
      it.p[1] = y;
      items.push_back(it);
    }
  }
  void InjectMoveToIfNeed() {
    if (last_move_to_index < 0) {
      float x = 0, y = 0;
      if (!verbs.empty()) {
        // a cubic end can never be a contour start, so this is always an exact point
        const LPoint& pt = points[static_cast<size_t>(~last_move_to_index)];
        x = pt.x;
        y = pt.y;
      }
      MoveTo(x, y);
    }
  }
  void Segment(uint8_t verb, const Point pts[4], int n_pts, float w) {
    InjectMoveToIfNeed();
    Item it{};
    it.verb = verb;
    it.w = w;
    for (int i = 0; i < n_pts; i++) {
      it.p[2 * i] = pts[i].x;
      it.p[2 * i + 1] = pts[i].y;
    }
    items.push_back(it);
    verbs.push_back(verb);
    // only the last point matters for chaining; interior control points never become starts
    points.push_back(LPoint{pts[n_pts - 1].x, pts[n_pts - 1].y, verb == 4});
  }
  void Close() {
    if (!verbs.empty() && verbs.back() != 5) {
      verbs.push_back(5);
      Item it{};
      it.verb = 5;
      items.push_back(it);
    }
    last_move_to_index ^= ~last_move_to_index >> (8 * sizeof(last_move_to_index) - 1);
  }
};

}  // namespace

void LowerPathToSegs(const Path& src, std::vector<skb_dl_seg>* out) {
  // one scratch path per thread, reused from call to call: its three vectors keep their capacity, so lowering a path
  // allocates nothing (a Canvas user's thread and the builder's outline threads each have their own)
  static thread_local LoweredPath scratch;
  LoweredPath& dst = scratch;
  dst.items.clear();
  dst.points.clear();
  dst.verbs.clear();
  dst.last_move_to_index = ~0;
  {
    Path::Iter iter{src, false};
    Point pts[4] = {};
    for (;;) {
      Path::Verb verb = iter.Next(pts);
      bool done = false;
      switch (verb) {
        case Path::Verb::kMove:
          dst.MoveTo(pts[0].x, pts[0].y);
          break;
        case Path::Verb::kLine:
          dst.Segment(1, pts, 2, 0.f);
          break;
        case Path::Verb::kQuad:
          dst.Segment(2, pts, 3, 0.f);
          break;
        case Path::Verb::kConic:
          dst.Segment(3, pts, 3, iter.ConicWeight());
          break;
        case Path::Verb::kCubic:
          dst.Segment(4, pts, 4, 0.f);
          break;
        case Path::Verb::kClose:
          dst.Close();
          break;
        case Path::Verb::kDone:
          done = true;
          break;
      }
      if (done) break;
    }
  }

  // PathEdgeIter traversal (src/graphic/path_priv.hpp:75-167): auto-close every contour.
  bool needs_close = false;
  LPoint move_pt{0, 0, false};
  LPoint last_pt{0, 0, false};
  size_t contour_first_seg = 0;
  bool contour_open = false;
  auto emit = [&](uint32_t type, const float* p, float w, const LPoint& start) {
    skb_dl_seg s{};
    s.type_flags = type | (start.cubic_end ? SKB_SEG_P0_FROM_PREV_CUBIC : 0u);
    s.w = w;
    if (p) std::memcpy(s.p, p, sizeof(s.p));
    s.start[0] = start.x;
    s.start[1] = start.y;
    out->push_back(s);
  };
  auto closeline = [&]() {
    float p[8] = {last_pt.x, last_pt.y, move_pt.x, move_pt.y, 0, 0, 0, 0};
    emit(SKB_SEG_CLOSE, p, 0.f, last_pt);
    needs_close = false;
    last_pt = move_pt;
  };
  auto end_contour = [&]() {
    if (contour_open && out->size() == contour_first_seg) {
      // a MoveTo that stays in the path without any segment: bounds only
      emit(SKB_SEG_POINT, nullptr, 0.f, move_pt);
    }
    contour_open = false;
  };
  size_t pi = 0;
  for (const auto& it : dst.items) {
    switch (it.verb) {
      case 0:
        if (needs_close) closeline();
        end_contour();
        move_pt = dst.points[pi++];
        last_pt = move_pt;
        contour_open = true;
        contour_first_seg = out->size();
        break;
      case 5:
        if (needs_close) closeline();
        break;
      default: {
        static const uint32_t kType[5] = {0, SKB_SEG_LINE, SKB_SEG_QUAD, SKB_SEG_CONIC, SKB_SEG_CUBIC};
        emit(kType[it.verb], it.p, it.w, last_pt);
        last_pt = dst.points[pi++];
        needs_close = true;
      } break;
    }
  }
  if (needs_close) closeline();
  end_contour();
}

// ---------------------------------------------------------------------------
// Canvas
// ---------------------------------------------------------------------------
CudaCanvas::CudaCanvas(skb::DlBuilder* builder, uint32_t surface, uint32_t width, uint32_t height)
    : Canvas(), builder_(builder), surface_(surface), width_(width), height_(height) {
  state_stack_.emplace_back(State());
}

void CudaCanvas::NoteUnsupported(const char* what) {
  if (parent_canvas_) {
    parent_canvas_->NoteUnsupported(what);
    return;
  }
  if (unsupported_.empty()) unsupported_ = what;
}

// SWCanvas::CurrentTransform (sw_canvas.hpp:179-182)
Matrix CudaCanvas::CurrentTransform() const {
  return Matrix::Translate(-global_offset_.x, -global_offset_.y) * GetTotalMatrix();
}

// SWCanvas::GetScanClipBounds (sw_canvas.hpp:157-161)
Rect CudaCanvas::ScanClipBounds() const {
  Rect clip_bounds = GetGlobalClipBounds();
  clip_bounds.Offset(-global_offset_.x, -global_offset_.y);
  return clip_bounds;
}

// SWCanvas::OnSave (sw_canvas.cc:679-686)
void CudaCanvas::OnSave() {
  if (PeekLayerStack()) {
    PeekLayerStack()->canvas->Save();
    return;
  }
  state_stack_.emplace_back(state_stack_.back());
}

// SWCanvas::OnRestore (sw_canvas.cc:688-708)
void CudaCanvas::OnRestore() {
  if (state_stack_.size() == 1) return;
  if (PeekLayerStack()) {
    CudaCanvas* sub_canvas = PeekLayerStack()->canvas.get();
    if (sub_canvas->state_stack_.size() > 1) {
      sub_canvas->Restore();
      return;
    }
  }
  if (state_stack_.back().has_layer) OnLayerRestore();
  state_stack_.pop_back();
}

void CudaCanvas::OnRestoreToCount(int saveCount) {
  if (saveCount < 1) return;
  while (state_stack_.size() > static_cast<size_t>(saveCount)) this->OnRestore();
}

void CudaCanvas::OnFlush() {}

// SWCanvas::OnClipRect (sw_canvas.cc:297-313): an intersecting rect under a
// scale/translate CTM only tightens the integer scan rectangle.
void CudaCanvas::OnClipRect(const Rect& rect, ClipOp op) {
  SKITY_TRACE_EVENT(CudaCanvas_OnClipRect);
  if (PeekLayerStack()) {
    PeekLayerStack()->canvas->ClipRect(rect, op);
    return;
  }
  if (op == ClipOp::kDifference || !CurrentTransform().OnlyScaleAndTranslate()) {
    Canvas::OnClipRect(rect, op);
    return;
  }
}

// SWCanvas::OnClipPath (sw_canvas.cc:315-336)
void CudaCanvas::OnClipPath(const Path& path, ClipOp op) {
  SKITY_TRACE_EVENT(CudaCanvas_OnClipPath);
  if (PeekLayerStack()) {
    PeekLayerStack()->canvas->ClipPath(path, op);
    return;
  }
  if (surface_ == kNoSurface) return;
  std::vector<skb_dl_seg> segs;
  LowerPathToSegs(path, &segs);
  skb_dl_op o{};
  o.kind = SKB_OP_CLIP;
  o.surface = surface_;
  o.path = builder_->AddPath(segs);
  o.clip_in = state_stack_.back().clip_id;
  o.clip_out = builder_->NewClipState();
  o.fill_type = path.GetFillType() == Path::PathFillType::kEvenOdd ? 1u : 0u;
  o.aux = op == ClipOp::kIntersect ? 1u : 0u;
  Matrix m = CurrentTransform();
  o.ctm[0] = m.GetScaleX();
  o.ctm[1] = m.GetSkewX();
  o.ctm[2] = m.GetTranslateX();
  o.ctm[3] = m.GetSkewY();
  o.ctm[4] = m.GetScaleY();
  o.ctm[5] = m.GetTranslateY();
  Rect cb = ScanClipBounds();
  o.clip_bounds[0] = cb.Left();
  o.clip_bounds[1] = cb.Top();
  o.clip_bounds[2] = cb.Right();
  o.clip_bounds[3] = cb.Bottom();
  builder_->AddOp(o);
  state_stack_.back().clip_id = o.clip_out;
}

void CudaCanvas::EmitFill(const Path& path, const Matrix& m, uint32_t paint_index) {
  if (surface_ == kNoSurface) return;
  if (m.HasPersp()) {
    NoteUnsupported("perspective CTM");
    return;
  }
  static thread_local std::vector<skb_dl_seg> segs;   // reused: AddPath copies what it holds
  segs.clear();
  LowerPathToSegs(path, &segs);
  if (segs.empty()) return;
  EmitFillOp(builder_->AddPath(segs), path.GetFillType(), m, paint_index);
}

// The fill op of a path that is already in the builder (possibly still being worked out: AddDeferredPath).
void CudaCanvas::EmitFillOp(uint32_t path_index, Path::PathFillType fill_type, const Matrix& m, uint32_t paint_index) {
  skb_dl_op o{};
  o.kind = SKB_OP_FILL;
  o.surface = surface_;
  o.path = path_index;
  o.paint = paint_index;
  o.clip_in = state_stack_.back().clip_id;
  o.fill_type = fill_type == Path::PathFillType::kEvenOdd ? 1u : 0u;
  o.ctm[0] = m.GetScaleX();
  o.ctm[1] = m.GetSkewX();
  o.ctm[2] = m.GetTranslateX();
  o.ctm[3] = m.GetSkewY();
  o.ctm[4] = m.GetScaleY();
  o.ctm[5] = m.GetTranslateY();
  Rect cb = ScanClipBounds();
  o.clip_bounds[0] = cb.Left();
  o.clip_bounds[1] = cb.Top();
  o.clip_bounds[2] = cb.Right();
  o.clip_bounds[3] = cb.Bottom();
  builder_->AddOp(o);
}

namespace {

// PointsToUnit (src/render/sw/sw_span_brush.cc:164-198), same Matrix calls in the same order.
Matrix PointsToUnit(const Shader::GradientInfo& info, Shader::GradientType type) {
  if (type == Shader::GradientType::kLinear) {
    Vec2 start = Vec2{(info.point[0])};
    Vec2 stop = Vec2{info.point[1]};
    Vec2 ss = stop - start;
    float length = ss.Length();
    float scale = length > 0 ? 1.0f / length : 0;
    Vec2 unit_ss = ss * scale;
    float sine = -unit_ss.y;
    float cosine = unit_ss.x;
    Matrix rotate;
    rotate.SetScaleX(cosine);
    rotate.SetSkewX(-sine);
    rotate.SetSkewY(sine);
    rotate.SetScaleY(cosine);
    return rotate * Matrix::Scale(scale, scale) * Matrix::Translate(-start.x, -start.y);
  } else if (type == Shader::GradientType::kRadial) {
    float radius = info.radius[0];
    Vec2 center = Vec2{info.point[0]};
    float scale = radius > 0 ? 1.0f / radius : 0;
    return Matrix::Scale(scale, scale) * Matrix::Translate(-center.x, -center.y);
  } else if (type == Shader::GradientType::kSweep) {
    Vec2 center = Vec2{info.point[0]};
    return Matrix::Translate(-center.x, -center.y);
  }
  return Matrix{};
}

// PointsToUnit(p0, p1) (sw_span_brush.cc:197-216): the unit-segment transform the conical brush uses
Matrix PointsToUnit2(const Point& p0, const Point& p1) {
  Shader::GradientInfo two{};
  two.point[0] = p0;
  two.point[1] = p1;
  return PointsToUnit(two, Shader::GradientType::kLinear);
}

void StoreAffine(const Matrix& m, float out[6]) {
  out[0] = m.GetScaleX();
  out[1] = m.GetSkewX();
  out[2] = m.GetTranslateX();
  out[3] = m.GetSkewY();
  out[4] = m.GetScaleY();
  out[5] = m.GetTranslateY();
}

// ComputeBoundsIfStroke (src/render/sw/sw_canvas.cc:135-144)
Rect ComputeBoundsIfStroke(Rect bounds, const Paint& paint) {
  if (paint.GetStyle() != Paint::kFill_Style) {
    float stroke_width = paint.GetStrokeWidth();
    bounds.SetLTRB(std::floor(bounds.Left() - stroke_width), std::floor(bounds.Top() - stroke_width),
                   std::floor(bounds.Right() + stroke_width), std::floor(bounds.Bottom() + stroke_width));
  }
  return bounds;
}

}  // namespace

// SWCanvas::GenerateBrush (sw_canvas.cc:727-795), not-drawing-layer branch.
// SWRenderTarget implements kClear..kScreen and kSoftLight; everything else falls back to kSrcOver
// (src/graphic/blend_mode.cc:129-133).  0 encodes the default.
static uint32_t EncodeBlend(BlendMode mode) {
  auto m = static_cast<int32_t>(mode);
  if (m < 0 || (m > static_cast<int32_t>(BlendMode::kScreen) && mode != BlendMode::kSoftLight)) return 0;
  return mode == BlendMode::kSrcOver ? 0u : static_cast<uint32_t>(m) + 1u;
}

// Encodes paint.GetColorFilter() as a block of the float pool (include/skb_dl.h, SKB_CF_*) and returns the bits to
// OR into skb_dl_paint::has_stops.  What each filter computes: src/effect/color_filter.cc:123-197.
uint32_t CudaCanvas::EncodeColorFilter(const Paint& paint) {
  auto filter = paint.GetColorFilter();
  if (!filter) return 0;
  const ColorFilterBase* base = As_CFB(filter.get());
  uint32_t blk[68] = {};
  uint32_t n_words = 16;
  switch (base->GetType()) {
    case ColorFilterType::kBlend: {
      auto* f = static_cast<const BlendColorFilter*>(base);
      blk[0] = SKB_CF_BLEND;
      uint32_t mode = EncodeBlend(f->GetBlendMode());  // same fall-back to kSrcOver as a draw's blend mode
      blk[1] = mode ? mode - 1 : static_cast<uint32_t>(BlendMode::kSrcOver);
      blk[2] = ColorToPMColor(f->GetColor());
    } break;
    case ColorFilterType::kMatrix: {
      auto* f = const_cast<MatrixColorFilter*>(static_cast<const MatrixColorFilter*>(base));
      auto [mul, add] = f->GetMatrix();  // column i of `mul` = coefficients of input channel i
      blk[0] = SKB_CF_MATRIX;
      int16_t m16[20];
      for (int r = 0; r < 4; r++) {
        for (int c = 0; c < 4; c++) m16[5 * r + c] = static_cast<int16_t>(mul.Get(r, c) * 255);
        m16[5 * r + 4] = static_cast<int16_t>(add[r] * 255);
      }
      std::memcpy(blk + 4, m16, sizeof(m16));
    } break;
    case ColorFilterType::kLinearToSRGBGamma:
    case ColorFilterType::kSRGBToLinearGamma: {
      // the per-channel table is read off the filter itself: an opaque grey goes through unchanged except for the look-up
      blk[0] = SKB_CF_TABLE;
      uint8_t* table = reinterpret_cast<uint8_t*>(blk + 4);
      for (uint32_t v = 0; v < 256; v++) table[v] = static_cast<uint8_t>(ColorGetR(filter->FilterColor(ColorSetARGB(255, v, v, v))));
      n_words = 68;
    } break;
    case ColorFilterType::kCompose:
      return 0;  // ComposeColorFilter::OnFilterColor returns its input (color_filter.cc:194-197)
    default:
      NoteUnsupported("colour filter type");
      return 0;
  }
  uint32_t off = builder_->AddWords(blk, n_words);
  return (off + 1) << 8;
}

uint32_t CudaCanvas::MakeBrush(const Paint& paint, bool stroke) {
  skb_dl_paint p{};
  p.global_alpha = 255;
  p.blend = EncodeBlend(paint.GetBlendMode());
  uint32_t cf_bits = 0;
  if (paint.GetColorFilter()) {
    cf_bits = EncodeColorFilter(paint);
  }
  p.has_stops = cf_bits;
  auto shader = paint.GetShader();
  if (shader) {
    Shader::GradientInfo info{};
    Shader::GradientType type = shader->AsGradient(&info);
    if (type == Shader::kLinear || type == Shader::kRadial || type == Shader::kSweep || type == Shader::kConical) {
      Matrix device_to_local;
      shader->GetLocalMatrix().Invert(&device_to_local);
      Matrix layer_to_local;
      CurrentTransform().Invert(&layer_to_local);
      device_to_local = device_to_local * layer_to_local;
      // the conical brush keeps device_to_local apart from its unit transforms (sw_span_brush.cc:393-394)
      Matrix ptu = type == Shader::kConical ? device_to_local : PointsToUnit(info, type) * device_to_local;
      StoreAffine(ptu, p.m);
      p.type = type == Shader::kLinear ? SKB_PAINT_LINEAR
                                       : (type == Shader::kRadial ? SKB_PAINT_RADIAL
                                                                  : (type == Shader::kSweep ? SKB_PAINT_SWEEP : SKB_PAINT_CONICAL));
      p.tile_mode = static_cast<uint32_t>(info.tile_mode);
      p.n_colors = static_cast<uint32_t>(info.colors.size());
      p.has_stops = cf_bits | (info.color_offsets.empty() ? 0u : 1u);
      std::vector<float> cols(4 * info.colors.size());
      for (size_t i = 0; i < info.colors.size(); i++) {
        cols[4 * i + 0] = info.colors[i].x;
        cols[4 * i + 1] = info.colors[i].y;
        cols[4 * i + 2] = info.colors[i].z;
        cols[4 * i + 3] = info.colors[i].w;
      }
      std::vector<float> offs(info.colors.size(), 0.f);
      for (size_t i = 0; i < info.color_offsets.size() && i < offs.size(); i++) offs[i] = info.color_offsets[i];
      p.stop_off = builder_->AddStops(cols.data(), offs.data(), p.n_colors);
      p.bias = info.radius[0];
      p.scale = info.radius[1];
      if (type == Shader::kConical) {
        // ConicalGradientColorBrush::OnPreBrush (sw_span_brush.cc:405-444), evaluated once here
        float e[16] = {};
        Point c0 = info.point[0], c1 = info.point[1];
        float r0 = info.radius[0], r1 = info.radius[1];
        float delta_center = (c1 - c0).Length();
        float delta_radius = std::abs(r1 - r0);
        if (!(r0 < 0 || r1 < 0)) {
          bool radial = delta_center < kNearlyZero;
          bool strip = delta_radius < kNearlyZero;
          if (radial) {
            e[0] = 4.f;  // concentric with equal radii: transparent
            if (!strip) {
              e[0] = 1.f;
              e[1] = c0.x;
              e[2] = c0.y;
              e[3] = 1.0 / delta_radius;
              e[4] = delta_radius < 0 ? -1.f : 1.f;
              e[5] = r0 / delta_radius;
            }
          } else if (strip) {
            e[0] = 2.f;
            e[6] = r0 / delta_center;
            StoreAffine(PointsToUnit2(c0, c1), e + 7);
          } else {
            bool swap_01 = r1 < kNearlyZero;
            if (swap_01) {
              std::swap(c0, c1);
              std::swap(r0, r1);
            }
            float f = r0 / (r0 - r1);
            Point cf = c0 * (1.f - f) + c1 * f;
            r1 = r1 / (c1 - cf).Length();
            e[0] = swap_01 ? 5.f : 3.f;
            StoreAffine(PointsToUnit2(cf, c1), e + 7);
            e[13] = r1;
            e[14] = r1 * r1;
            e[15] = f;
          }
        }
        builder_->AddFloats(e, 16);
      }
      return builder_->AddPaint(p);
    }
    if (const auto* image_ptr = shader->AsImage()) {
      // GenerateBrush, image branch (sw_canvas.cc:755-787) + PixmapBrush (sw_span_brush.cc:555-579)
      const auto& image = *image_ptr;
      const std::shared_ptr<Pixmap>* pm = image ? image->GetPixmap() : nullptr;
      if (!pm || !*pm || (*pm)->Width() == 0 || (*pm)->Height() == 0) {
        NoteUnsupported("image shader without CPU pixels (texture-backed image)");
      } else {
        auto pixmap_shader = std::static_pointer_cast<PixmapShader>(shader);
        const std::shared_ptr<Pixmap>& pixmap = *pm;
        uint32_t iw = pixmap->Width(), ih = pixmap->Height();
        // the image is identified by its pixmap and the bytes it holds right now; only a new (pixmap, content) pair
        // pays the conversion to the Colors Bitmap::GetPixel hands the sampler, whatever the pixmap's colour type
        const uint64_t content = skb::DlBuilder::HashBytes(pixmap->Addr(), pixmap->RowBytes() * static_cast<size_t>(ih));
        uint32_t sid = builder_->FindImageSurface(pixmap.get(), content, iw, ih);
        if (sid == 0) {
          std::vector<uint8_t> rgba(static_cast<size_t>(iw) * ih * 4);
          Bitmap bm(pixmap, true);
          for (uint32_t yy = 0; yy < ih; yy++) {
            for (uint32_t xx = 0; xx < iw; xx++) {
              Color c = bm.GetPixel(xx, yy);
              uint8_t* d = &rgba[(static_cast<size_t>(yy) * iw + xx) * 4];
              d[0] = ColorGetR(c);
              d[1] = ColorGetG(c);
              d[2] = ColorGetB(c);
              d[3] = ColorGetA(c);
            }
          }
          sid = builder_->AddImageSurface(pixmap.get(), content, pixmap, iw, ih, rgba.data());
        }
        Matrix inverse;
        shader->GetLocalMatrix().Invert(&inverse);
        Matrix matrix = Matrix::Scale(1.f / iw, 1.f / ih) * inverse;
        Matrix layer_to_local;
        CurrentTransform().Invert(&layer_to_local);
        matrix = matrix * layer_to_local;
        StoreAffine(matrix, p.m);
        const SamplingOptions& sampling = *pixmap_shader->GetSamplingOptions();
        const bool linear = sampling.UseCubic() || sampling.filter == FilterMode::kLinear;
        p.type = SKB_PAINT_IMAGE;
        p.tile_mode = static_cast<uint32_t>(pixmap_shader->GetXTileMode()) |
                      (static_cast<uint32_t>(pixmap_shader->GetYTileMode()) << 4) | SKB_PAINT_IMAGE_YMODE |
                      (linear ? SKB_PAINT_IMAGE_LINEAR : 0u) |
                      (pixmap->GetAlphaType() == kUnpremul_AlphaType ? SKB_PAINT_IMAGE_UNPREMUL : 0u);
        p.image_surface = sid;
        // SWSpanBrush converts the float alpha to its byte with a plain cast (sw_span_brush.hpp:29-35)
        p.global_alpha = static_cast<uint8_t>(255 * paint.GetAlphaF());
        return builder_->AddPaint(p);
      }
    }
  }
  Color4f color = stroke ? paint.GetStrokeColor() : paint.GetFillColor();
  p.type = SKB_PAINT_SOLID;
  p.color[0] = color.r;
  p.color[1] = color.g;
  p.color[2] = color.b;
  p.color[3] = color.a;
  return builder_->AddPaint(p);
}

void CudaCanvas::FillPath(const Path& path, const Paint& paint, bool stroke) {
  EmitFill(path, CurrentTransform(), MakeBrush(paint, stroke));
}

// SWCanvas::OnDrawPath (sw_canvas.cc:357-411)
void CudaCanvas::OnDrawPath(const Path& path, const Paint& paint) {
  SKITY_TRACE_EVENT(CudaCanvas_OnDrawPath);
  if (PeekLayerStack()) {
    PeekLayerStack()->canvas->DrawPath(path, paint);
    return;
  }
  if (paint.GetMaskFilter() || paint.GetImageFilter()) {
    HandleFilter(path, paint);
    return;
  }

  bool need_fill = paint.GetStyle() != Paint::kStroke_Style;
  bool need_stroke = paint.GetStyle() != Paint::kFill_Style;

  auto draw_fill = [&]() {
    Path temp;
    if (paint.GetPathEffect() && paint.GetPathEffect()->FilterPath(&temp, path, false, paint)) {
      FillPath(temp, paint, false);
    } else {
      FillPath(path, paint, false);
    }
  };
  // The outline (the reference's own Stroke, sw_canvas.cc:388-401) is worked out on one of the display-list builder's
  // threads from copies of the path and the paint; the fill op that draws it takes its place in the list right away.
  auto draw_stroke = [&]() {
    if (surface_ == kNoSurface) return;
    const Matrix m = CurrentTransform();
    if (m.HasPersp()) {
      NoteUnsupported("perspective CTM");
      return;
    }
    const uint32_t paint_index = MakeBrush(paint, true);
    const uint32_t path_index = builder_->AddDeferredPath([path, paint](std::vector<skb_dl_seg>* out) {
      Stroke stroke(paint);
      Path temp;
      Path quad;
      Path outline;
      if (paint.GetPathEffect() && paint.GetPathEffect()->FilterPath(&temp, path, true, paint)) {
        stroke.QuadPath(temp, &quad);
        stroke.StrokePath(quad, &outline);
      } else {
        stroke.QuadPath(path, &quad);
        stroke.StrokePath(quad, &outline);
      }
      LowerPathToSegs(outline, out);
    });
    EmitFillOp(path_index, Path::PathFillType::kWinding, m, paint_index);
  };
  // DrawFillStrokeInPaintOrder (src/render/paint_order.hpp:12-31)
  if (paint.GetStyle() == Paint::kStrokeThenFill_Style) {
    if (need_stroke) draw_stroke();
    if (need_fill) draw_fill();
  } else {
    if (need_fill) draw_fill();
    if (need_stroke) draw_stroke();
  }
}

// SWCanvas::OnDrawPaint (sw_canvas.cc:413-439): the whole bitmap, identity transform, no scan clip.
void CudaCanvas::OnDrawPaint(const Paint& paint) {
  SKITY_TRACE_EVENT(CudaCanvas_OnDrawPaint);
  if (PeekLayerStack()) {
    PeekLayerStack()->canvas->DrawPaint(paint);
    return;
  }
  if (surface_ == kNoSurface) return;
  Rect bounds = Rect::MakeWH(Width(), Height());
  Path path;
  path.AddRect(bounds);
  std::vector<skb_dl_seg> segs;
  LowerPathToSegs(path, &segs);
  skb_dl_op o{};
  o.kind = SKB_OP_FILL;
  o.surface = surface_;
  o.path = builder_->AddPath(segs);
  o.paint = MakeBrush(paint, false);
  o.clip_in = state_stack_.back().clip_id;
  o.fill_type = 0;
  StoreAffine(Matrix{}, o.ctm);
  o.clip_bounds[0] = -1E9F;  // SWRaster::kCullRect (sw_raster.hpp:84)
  o.clip_bounds[1] = -1E9F;
  o.clip_bounds[2] = 1E9F;
  o.clip_bounds[3] = 1E9F;
  builder_->AddOp(o);
}

// SWCanvas::HandleFilter + MaskFilterOnFilter(kNormal) + ImageFilterBase::BlurBitmapToCanvas
// (sw_canvas.cc:797-826, mask_filter.cc:51-60, image_filter.cc:33-41,184-194).
// SWCanvas::HandleFilter (sw_canvas.cc:797-826) with MaskFilterOnFilter (src/effect/mask_filter.cc:51-103),
// BlurImageFilter::OnFilter and DropShadowImageFilter::OnFilter (src/effect/image_filter.cc:196-238): the
// path is drawn into an offscreen surface, blurred into a second one, post-processed per pixel for the
// blur styles / the shadow colour, and composited back as an image.
void CudaCanvas::HandleFilter(const Path& path, const Paint& paint) {
  SKITY_TRACE_EVENT(CudaCanvas_HandleFilter);
  HandleFilterOf(path.GetBounds(), paint, [&](CudaCanvas& temp_canvas, const Paint& work_paint) { temp_canvas.DrawPath(path, work_paint); });
}

// `draw_source` draws what is to be filtered (a path, or an image that lives on the device) into the temporary canvas.
void CudaCanvas::HandleFilterOf(const Rect& source_bounds, const Paint& paint,
                                const std::function<void(CudaCanvas&, const Paint&)>& draw_source) {
  Paint work_paint = paint;
  work_paint.SetMaskFilter(nullptr);
  work_paint.SetImageFilter(nullptr);

  auto mask_filter = paint.GetMaskFilter();
  auto image_filter = As_IFB(paint.GetImageFilter().get());
  if (!mask_filter) {
    auto type = image_filter->GetType();
    if (type != ImageFilterType::kBlur && type != ImageFilterType::kDropShadow && type != ImageFilterType::kDilate &&
        type != ImageFilterType::kErode) {
      NoteUnsupported("image filter other than Blur / DropShadow / Dilate / Erode");  // the SW backend has no OnFilter for them either
      return;
    }
  }
  Rect bounds = ComputeBoundsIfStroke(source_bounds, paint);
  float radius_x = mask_filter ? mask_filter->GetBlurRadius() : image_filter->GetRadiusX();
  float radius_y = mask_filter ? mask_filter->GetBlurRadius() : image_filter->GetRadiusY();
  Rect fb = Rect::MakeLTRB(std::floor(bounds.Left() - radius_x), std::floor(bounds.Top() - radius_y),
                           std::ceil(bounds.Right() + radius_x), std::ceil(bounds.Bottom() + radius_y));
  uint32_t w = static_cast<uint32_t>(fb.Width());
  uint32_t h = static_cast<uint32_t>(fb.Height());
  if (w == 0 || h == 0) return;  // the reference dereferences a null temp canvas here

  uint32_t temp = builder_->AddSurface(w, h);
  {
    CudaCanvas temp_canvas(builder_, temp, w, h);
    temp_canvas.Translate(-fb.Left(), -fb.Top());
    draw_source(temp_canvas, work_paint);
    if (!temp_canvas.Unsupported().empty()) NoteUnsupported(temp_canvas.Unsupported().c_str());
  }
  uint32_t blurred = builder_->AddSurface(w, h);
  skb_dl_op b{};
  b.kind = SKB_OP_BLUR;
  b.surface = blurred;
  b.aux = temp;
  b.clip_bounds[0] = std::round(std::max(radius_x, radius_y));
  if (mask_filter) {
    switch (mask_filter->GetBlurStyle()) {
      case BlurStyle::kSolid: b.fill_type = 2; break;
      case BlurStyle::kOuter: b.fill_type = 3; break;
      case BlurStyle::kInner: b.fill_type = 4; break;
      default: b.fill_type = 0; break;
    }
    builder_->AddOp(b);
    DrawSurfaceImage(blurred, w, h, fb, work_paint, false);
    return;
  }
  if (image_filter->GetType() == ImageFilterType::kBlur) {
    builder_->AddOp(b);
    DrawSurfaceImage(blurred, w, h, fb, work_paint, false);
    return;
  }
  if (image_filter->GetType() == ImageFilterType::kDilate || image_filter->GetType() == ImageFilterType::kErode) {
    // MorphologyImageFilter::OnFilter (image_filter.cc:342-385): the result bitmap is of the default (unpremultiplied) type
    b.fill_type = image_filter->GetType() == ImageFilterType::kDilate ? 6 : 7;
    b.clip_bounds[0] = radius_x;
    b.clip_bounds[1] = radius_y;
    builder_->AddOp(b);
    DrawSurfaceImage(blurred, w, h, fb, work_paint, true);
    return;
  }
  // drop shadow: the blurred alpha tinted with the (unpremultiplied) shadow colour, offset, then the shape
  b.fill_type = 5;
  b.paint = image_filter->GetColor();
  builder_->AddOp(b);
  Rect offset_bounds = fb;
  offset_bounds.Offset(image_filter->GetOffsetX(), image_filter->GetOffsetY());
  DrawSurfaceImage(blurred, w, h, offset_bounds, work_paint, true);
  DrawSurfaceImage(temp, w, h, fb, work_paint, false);
}

// Canvas::DrawImage(image, rect, paint) -> SWCanvas::OnDrawImageRect (sw_canvas.cc:641-677)
// -> GenerateBrush image branch (sw_canvas.cc:755-787), for an image that lives on the device.
void CudaCanvas::DrawSurfaceImage(uint32_t src_surface, uint32_t iw, uint32_t ih, const Rect& dst,
                                  const Paint& paint, bool unpremul, const Matrix* shader_local) {
  Rect src = Rect::MakeWH(iw, ih);
  if (src.Width() == 0 || src.Height() == 0 || dst.Width() == 0 || dst.Height() == 0) return;
  if (PeekLayerStack()) {  // OnDrawImageRect ends in OnDrawPath, which an open layer takes over (sw_canvas.cc:360-363)
    PeekLayerStack()->canvas->DrawSurfaceImage(src_surface, iw, ih, dst, paint, unpremul, shader_local);
    return;
  }
  Path path;
  path.AddRect(dst);
  // the image shader's local matrix as OnDrawImageRect builds it (sw_canvas.cc:657-667) — unless the caller is a
  // filter's temporary canvas replaying a shader that was built elsewhere
  Matrix local_matrix;
  if (shader_local) {
    local_matrix = *shader_local;
  } else if (IsDrawingLayer()) {
    local_matrix = Matrix::Scale(1.f / src.Width(), 1.f / src.Height()) * Matrix::Translate(-src.Left(), -src.Top());
  } else {
    local_matrix = Matrix::Translate(dst.Left(), dst.Top()) *
                   Matrix::Scale(dst.Width() / src.Width(), dst.Height() / src.Height()) *
                   Matrix::Translate(-src.Left(), -src.Top());
  }
  if (paint.GetMaskFilter() || paint.GetImageFilter()) {
    // OnDrawImageRect -> OnDrawPath(rect, image shader + filter) -> HandleFilter (sw_canvas.cc:656-676,365-369): the image
    // is resampled into the temporary canvas, filtered, and the result drawn back
    Paint fill_paint = paint;
    fill_paint.SetStyle(Paint::kFill_Style);
    HandleFilterOf(path.GetBounds(), fill_paint, [&](CudaCanvas& temp_canvas, const Paint& work_paint) {
      temp_canvas.DrawSurfaceImage(src_surface, iw, ih, dst, work_paint, unpremul, &local_matrix);
    });
    return;
  }
  Matrix inverse;
  local_matrix.Invert(&inverse);
  Matrix matrix = Matrix::Scale(1.f / iw, 1.f / ih) * inverse;
  if (IsDrawingLayer()) {  // (a filter's temporary canvas is a canvas of its own: never "drawing a layer")
    // GenerateBrush maps the raster bounds of the drawn rectangle onto the layer (sw_canvas.cc:772-776);
    // the bounds are SWRaster::RastePath's (sw_raster.cc:737-745)
    Paint plain;
    Stroke stroke(plain);
    Path quad;
    stroke.QuadPath(path, &quad);
    Rect sb = quad.CopyWithMatrix(Matrix(CurrentTransform())).GetBounds();
    Rect bounds = Rect::MakeLTRB(std::floor(sb.Left()), std::floor(sb.Top()), std::ceil(sb.Right()), std::ceil(sb.Bottom()));
    matrix = matrix * Matrix::Scale(1.0f / bounds.Width(), 1.0f / bounds.Height()) *
             Matrix::Translate(-bounds.Left(), -bounds.Top());
  } else {
    Matrix layer_to_local;
    CurrentTransform().Invert(&layer_to_local);
    matrix = matrix * layer_to_local;
  }

  skb_dl_paint p{};
  p.type = SKB_PAINT_IMAGE;
  p.tile_mode = 3u | (unpremul ? SKB_PAINT_IMAGE_UNPREMUL : 0u);
  StoreAffine(matrix, p.m);
  p.image_surface = src_surface;
  // work_paint.SetStyle(kFill) precedes GetAlphaF() in the reference (sw_canvas.cc:656,784)
  p.global_alpha = static_cast<uint8_t>(255 * paint.GetFillColor().a);
  p.blend = EncodeBlend(paint.GetBlendMode());
  if (paint.GetColorFilter()) {
    p.has_stops = EncodeColorFilter(paint);
  }
  uint32_t paint_index = builder_->AddPaint(p);
  EmitFill(path, CurrentTransform(), paint_index);
}

// SWCanvas::OnSaveLayer + GenerateLayer + LayerState::Init (sw_canvas.cc:267-280,441-484,880-889):
// an offscreen surface as large as the layer's device-space bounds, drawn into by a sub-canvas that
// shares this canvas's CTM stack and global clip, shifted by the layer's device-space origin.
void CudaCanvas::OnSaveLayer(const Rect& bounds, const Paint& paint) {
  SKITY_TRACE_EVENT(CudaCanvas_OnSaveLayer);
  state_stack_.emplace_back(state_stack_.back());
  if (PeekLayerStack()) PeekLayerStack()->canvas->OnSave();
  state_stack_.back().has_layer = true;

  Paint work_paint{paint};
  work_paint.SetStyle(Paint::kFill_Style);
  auto layer_bounds = work_paint.ComputeFastBounds(bounds);

  CudaCanvas* target_canvas = this;
  if (PeekLayerStack()) target_canvas = PeekLayerStack()->canvas.get();

  auto canvas_matrix = target_canvas->CurrentTransform();
  Rect device_bounds = Rect::MakeWH(target_canvas->Width(), target_canvas->Height());
  Rect scan_clip_bounds = target_canvas->ScanClipBounds();
  if (!device_bounds.Intersect(scan_clip_bounds)) device_bounds.SetEmpty();

  Matrix device_to_local;
  if (canvas_matrix.Invert(&device_to_local)) {
    Rect local_clip_bounds;
    device_to_local.MapRect(&local_clip_bounds, device_bounds);
    if (!layer_bounds.Intersect(local_clip_bounds)) layer_bounds.SetEmpty();
  }
  Rect rel_bounds{};
  canvas_matrix.MapRect(&rel_bounds, layer_bounds);

  auto layer = std::make_unique<LayerState>();
  layer->rel_bounds = rel_bounds;
  layer->log_bounds = layer_bounds;
  layer->width = static_cast<uint32_t>(std::ceil(rel_bounds.Width()));
  layer->height = static_cast<uint32_t>(std::ceil(rel_bounds.Height()));
  layer->surface = (layer->width && layer->height) ? builder_->AddSurface(layer->width, layer->height) : kNoSurface;
  layer->canvas = std::make_unique<CudaCanvas>(builder_, layer->surface, layer->width, layer->height);
  layer->canvas->SetTracingCanvasState(false);
  layer->canvas->parent_canvas_ = this;
  layer->canvas->global_offset_ = target_canvas->global_offset_ + Vec2{rel_bounds.Left(), rel_bounds.Top()};
  layer->paint = paint;
  layer_stack_.emplace_back(std::move(layer));
}

// SWCanvas::OnLayerRestore (sw_canvas.cc:891-902): the layer is drawn back as an image with the layer paint
void CudaCanvas::OnLayerRestore() {
  auto layer = std::move(layer_stack_.back());
  layer_stack_.pop_back();
  if (PeekLayerStack()) PeekLayerStack()->canvas->OnRestore();
  SetDrawingLayer(layer->paint.GetMaskFilter() == nullptr);
  if (layer->surface != kNoSurface) {
    DrawSurfaceImage(layer->surface, layer->width, layer->height, layer->log_bounds, layer->paint, false);
  }
  SetDrawingLayer(false);
}

void CudaCanvas::OnDrawBlob(const TextBlob*, float, float, Paint const&) { NoteUnsupported("text"); }

void CudaCanvas::OnDrawGlyphs(uint32_t, const GlyphID*, const float*, const float*, const Font&,
                              const Paint&) {
  NoteUnsupported("text");
}

// SWCanvas::OnDrawImageRect (sw_canvas.cc:641-677): the image becomes a decal shader over the destination rectangle
void CudaCanvas::OnDrawImageRect(std::shared_ptr<Image> image, const Rect& src, const Rect& dst, const SamplingOptions& sampling,
                                 Paint const* paint) {
  SKITY_TRACE_EVENT(CudaCanvas_OnDrawImageRect);
  if (!image) return;
  if (src.Width() == 0 || src.Height() == 0 || dst.Width() == 0 || dst.Height() == 0) return;
  Paint work_paint = (paint == nullptr) ? Paint() : *paint;
  work_paint.SetStyle(Paint::kFill_Style);
  Matrix local_matrix = Matrix::Translate(dst.Left(), dst.Top()) *
                        Matrix::Scale(dst.Width() / src.Width(), dst.Height() / src.Height()) *
                        Matrix::Translate(-src.Left(), -src.Top());
  auto shader = Shader::MakeShader(std::move(image), sampling, TileMode::kDecal, TileMode::kDecal, local_matrix);
  work_paint.SetShader(std::move(shader));
  Path path;
  path.AddRect(dst);
  this->OnDrawPath(path, work_paint);
}

}  // namespace skity
