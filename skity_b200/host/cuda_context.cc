// CudaGPUContext / CudaGPUSurface: the skity::GPUContext and skity::GPUSurface subclasses of the
// B200 backend.  Both public classes are pure-virtual interfaces (include/skity/gpu/gpu_context.hpp:69-,
// gpu_surface.hpp:67-130), so the backend subclasses them directly and needs none of the
// reference's GPUContextImpl / GPUDevice / WGSL machinery.  Everything device-side is reached
// through the C ABI of include/skb.h.
#include <cstring>
#include <skity/io/pixmap.hpp>
#include <string>

#include "include/skb.h"
#include "skity_b200/host/cuda_canvas.hpp"
#include "skity_b200/host/gpu_context_cuda.hpp"

namespace skity {

namespace {

class CudaGPUContext;

class CudaGPUSurface : public GPUSurface {
 public:
  CudaGPUSurface(GPUContext* ctx, skb_surface surface, uint32_t w, uint32_t h, float scale)
      : ctx_(ctx), surface_(surface), width_(w), height_(h), content_scale_(scale) {}
  ~CudaGPUSurface() override {
    skb_surface_destroy(surface_);
    skb_host_free(dl_buf_);
  }

  GPUBackendType GetBackendType() const override { return kGPUBackendTypeCUDA; }
  uint32_t GetWidth() const override { return width_; }
  uint32_t GetHeight() const override { return height_; }
  float ContentScale() const override { return content_scale_; }

  // The canvas is owned by the surface and valid until Flush (gpu_surface.hpp:96-104).
  Canvas* LockCanvas(bool clear) override {
    builder_.Reset(width_, height_);
    canvas_ = std::make_unique<CudaCanvas>(&builder_, 0u, width_, height_);
    if (skb_frame_begin(surface_, clear ? 1 : 0) != SKB_SUCCESS) Report();
    return canvas_.get();
  }

  void Flush() override {
    if (!canvas_) return;
    canvas_->Flush();
    if (!canvas_->Unsupported().empty()) {
      std::string msg = "skity-b200: dropped draws using " + canvas_->Unsupported();
      ctx_->TriggerErrorCallback(GPUError::kGPUError, msg.c_str());
    }
    // The frame's display list is laid out straight into a page-locked buffer the surface keeps (grown geometrically):
    // the upload is then one asynchronous copy at PCIe speed.  skb_frame_flush returns after the device has consumed
    // the list's tables, so the buffer is free again when the next frame is flushed.
    builder_.Finish();   // the stroke outlines still being worked out on the builder's threads
    const skb::DlBuilder::Layout layout = builder_.MakeLayout();
    const size_t need = layout.h.total_bytes;
    if (need > dl_cap_) {
      skb_host_free(dl_buf_);
      dl_buf_ = nullptr;
      dl_cap_ = 0;
      const size_t want = need + need / 2 + (static_cast<size_t>(1) << 20);
      void* p = nullptr;
      if (skb_host_alloc(want, &p) != SKB_SUCCESS) {
        Report();
        canvas_.reset();
        return;
      }
      dl_buf_ = p;
      dl_cap_ = want;
    }
    builder_.SerializeInto(layout, static_cast<uint8_t*>(dl_buf_));
    if (skb_frame_encode(surface_, dl_buf_, need) != SKB_SUCCESS || skb_frame_flush(surface_) != SKB_SUCCESS) {
      Report();
    }
    canvas_.reset();
  }

  std::shared_ptr<Pixmap> ReadPixels(const Rect& rect) override {
    Rect r = rect;
    if (!r.Intersect(Rect::MakeWH(width_, height_))) return nullptr;
    uint32_t x = static_cast<uint32_t>(r.Left()), y = static_cast<uint32_t>(r.Top());
    uint32_t w = static_cast<uint32_t>(r.Width()), h = static_cast<uint32_t>(r.Height());
    if (w == 0 || h == 0) return nullptr;
    auto pixmap = std::make_shared<Pixmap>(w, h, AlphaType::kPremul_AlphaType, ColorType::kRGBA);
    void* dst = pixmap->WritableAddr();
    if (!dst) return nullptr;
    if (skb_surface_read_pixels(surface_, x, y, w, h, dst, pixmap->RowBytes()) != SKB_SUCCESS) {
      Report();
      return nullptr;
    }
    return pixmap;
  }

  skb_surface Handle() const { return surface_; }

 private:
  void Report() { ctx_->TriggerErrorCallback(GPUError::kGPUError, skb_get_last_error_string()); }

  GPUContext* ctx_;
  skb_surface surface_;
  uint32_t width_, height_;
  float content_scale_;
  skb::DlBuilder builder_;
  std::unique_ptr<CudaCanvas> canvas_;
  void* dl_buf_ = nullptr;   // page-locked (skb_host_alloc)
  size_t dl_cap_ = 0;
};

class CudaGPUContext : public GPUContext {
 public:
  explicit CudaGPUContext(skb_device device) : device_(device) {}
  ~CudaGPUContext() override { skb_device_destroy(device_); }

  GPUBackendType GetBackendType() const override { return kGPUBackendTypeCUDA; }

  std::unique_ptr<GPUSurface> CreateSurface(GPUSurfaceDescriptor* desc) override {
    if (!desc || desc->backend != kGPUBackendTypeCUDA || desc->width == 0 || desc->height == 0) return nullptr;
    auto* cd = static_cast<GPUSurfaceDescriptorCuda*>(desc);
    skb_surface s = nullptr;
    if (skb_surface_create(device_, desc->width, desc->height, &s) != SKB_SUCCESS) {
      TriggerErrorCallback(GPUError::kGPUError, skb_get_last_error_string());
      return nullptr;
    }
    if (cd->band_y1 > cd->band_y0 && skb_surface_set_band(s, cd->band_y0, cd->band_y1) != SKB_SUCCESS) {
      TriggerErrorCallback(GPUError::kGPUError, skb_get_last_error_string());
      skb_surface_destroy(s);
      return nullptr;
    }
    return std::make_unique<CudaGPUSurface>(this, s, desc->width, desc->height, desc->content_scale);
  }

  // Presentation, textures, render targets and semaphores belong to the windowed graphics-API
  // backends; like GPUContextImpl does for what it lacks (src/gpu/gpu_context_impl.hpp:27-31,67-70)
  // they answer null here.
  std::unique_ptr<GPUPresenter> CreatePresenter(GPUPresenterDescriptor*) override { return nullptr; }
  std::shared_ptr<Texture> CreateTexture(TextureFormat, uint32_t, uint32_t, AlphaType) override { return nullptr; }
  std::shared_ptr<Texture> CreateTextureWithDesc(const TextureDescriptor*) override { return nullptr; }
  std::shared_ptr<Texture> WrapTexture(GPUBackendTextureInfo*, ReleaseCallback, ReleaseUserData) override {
    return nullptr;
  }
  std::unique_ptr<GPURenderTarget> CreateRenderTarget(const GPURenderTargetDescriptor&) override { return nullptr; }
  std::shared_ptr<Image> MakeSnapshot(std::unique_ptr<GPURenderTarget>) override { return nullptr; }
  void SetResourceCacheLimit(size_t) override {}
  std::shared_ptr<GPUSemaphore> CreateSemaphore() override { return nullptr; }
  void ImportSemaphore(GPUSemaphore*, const GPUSemaphoreImportInfo&) override {}

 private:
  skb_device device_;
};

}  // namespace

std::unique_ptr<GPUContext> CudaContextCreate(const CudaContextDesc* desc) {
  skb_device dev = nullptr;
  if (skb_device_create(desc ? desc->device_ordinal : 0, &dev) != SKB_SUCCESS) return nullptr;
  return std::make_unique<CudaGPUContext>(dev);
}

}  // namespace skity

// C entry point for harnesses: replay an SKSC scene through the COMPLETE plug-in path
// (CudaContextCreate -> CreateSurface -> LockCanvas -> Canvas calls -> Flush -> ReadPixels).
#include "skity_b200/host/scene_player.hpp"

extern "C" int skbh_render_scene_cuda(const uint8_t* scene, size_t n, int device_ordinal, uint8_t* out_rgba,
                                      char* err, size_t err_cap) {
  auto fail = [&](const char* m, int rc) {
    if (err && err_cap) {
      std::strncpy(err, m, err_cap - 1);
      err[err_cap - 1] = 0;
    }
    return rc;
  };
  if (n < sizeof(skb_scene::Header)) return fail("short scene", -1);
  skb_scene::Header h;
  std::memcpy(&h, scene, sizeof(h));
  if (h.magic != skb_scene::kMagic) return fail("bad magic", -1);
  skity::CudaContextDesc cd;
  cd.device_ordinal = device_ordinal;
  auto ctx = skity::CudaContextCreate(&cd);
  if (!ctx) return fail(skb_get_last_error_string(), -2);
  std::string cb_msg;
  ctx->SetErrorCallback(
      [](skity::GPUError, const char* message, void* user) { *static_cast<std::string*>(user) = message ? message : ""; },
      &cb_msg);
  skity::GPUSurfaceDescriptorCuda sd;
  sd.backend = skity::kGPUBackendTypeCUDA;
  sd.width = h.width;
  sd.height = h.height;
  auto surf = ctx->CreateSurface(&sd);
  if (!surf) return fail(cb_msg.c_str(), -3);
  skity::Canvas* canvas = surf->LockCanvas(true);
  int rc = skb_scene::Play(scene, n, canvas);
  if (rc != 0) return fail("malformed scene", rc);
  surf->Flush();
  if (!cb_msg.empty()) return fail(cb_msg.c_str(), -4);
  auto pm = surf->ReadPixels(skity::Rect::MakeWH(h.width, h.height));
  if (!pm) return fail(cb_msg.c_str(), -5);
  for (uint32_t y = 0; y < h.height; y++) {
    std::memcpy(out_rgba + static_cast<size_t>(y) * h.width * 4,
                static_cast<const uint8_t*>(pm->Addr()) + static_cast<size_t>(y) * pm->RowBytes(),
                static_cast<size_t>(h.width) * 4);
  }
  return 0;
}

// The same path over several frames on ONE context and surface, the way an application draws: per frame
// LockCanvas(true) -> Canvas calls -> Flush -> ReadPixels.  ms_out[0..3] receive the mean wall-clock milliseconds per
// frame of the frames after the first: whole frame, the Canvas calls (host encode), Flush (upload + device work is
// asynchronous: mostly the upload), ReadPixels (waits for the device, copies back).  out_rgba: the last frame.
#include <chrono>
extern "C" int skbh_render_scene_cuda_frames(const uint8_t* scene, size_t n, int device_ordinal, int frames, uint8_t* out_rgba,
                                             double* ms_out, char* err, size_t err_cap) {
  auto fail = [&](const char* m, int rc) {
    if (err && err_cap) {
      std::strncpy(err, m, err_cap - 1);
      err[err_cap - 1] = 0;
    }
    return rc;
  };
  if (n < sizeof(skb_scene::Header) || frames < 1) return fail("short scene", -1);
  skb_scene::Header h;
  std::memcpy(&h, scene, sizeof(h));
  if (h.magic != skb_scene::kMagic) return fail("bad magic", -1);
  skity::CudaContextDesc cd;
  cd.device_ordinal = device_ordinal;
  auto ctx = skity::CudaContextCreate(&cd);
  if (!ctx) return fail(skb_get_last_error_string(), -2);
  std::string cb_msg;
  ctx->SetErrorCallback(
      [](skity::GPUError, const char* message, void* user) { *static_cast<std::string*>(user) = message ? message : ""; },
      &cb_msg);
  skity::GPUSurfaceDescriptorCuda sd;
  sd.backend = skity::kGPUBackendTypeCUDA;
  sd.width = h.width;
  sd.height = h.height;
  auto surf = ctx->CreateSurface(&sd);
  if (!surf) return fail(cb_msg.c_str(), -3);
  using clk = std::chrono::steady_clock;
  auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  double acc[4] = {0, 0, 0, 0};
  for (int f = 0; f < frames; f++) {
    const auto t0 = clk::now();
    skity::Canvas* canvas = surf->LockCanvas(true);
    int rc = skb_scene::Play(scene, n, canvas);
    if (rc != 0) return fail("malformed scene", rc);
    const auto t1 = clk::now();
    surf->Flush();
    if (!cb_msg.empty()) return fail(cb_msg.c_str(), -4);
    const auto t2 = clk::now();
    auto pm = surf->ReadPixels(skity::Rect::MakeWH(h.width, h.height));
    if (!pm) return fail(cb_msg.c_str(), -5);
    const auto t3 = clk::now();
    if (f > 0 || frames == 1) {
      acc[0] += ms(t0, t3);
      acc[1] += ms(t0, t1);
      acc[2] += ms(t1, t2);
      acc[3] += ms(t2, t3);
    }
    if (f == frames - 1 && out_rgba) {
      for (uint32_t y = 0; y < h.height; y++) {
        std::memcpy(out_rgba + static_cast<size_t>(y) * h.width * 4,
                    static_cast<const uint8_t*>(pm->Addr()) + static_cast<size_t>(y) * pm->RowBytes(),
                    static_cast<size_t>(h.width) * 4);
      }
    }
  }
  const double cnt = frames > 1 ? frames - 1 : 1;
  if (ms_out)
    for (int i = 0; i < 4; i++) ms_out[i] = acc[i] / cnt;
  return 0;
}
