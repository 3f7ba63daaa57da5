// TEST INFRASTRUCTURE — not product code.
//
// C entry points over the UNMODIFIED reference software backend, compiled from
// the sources where they lie under /root/reference by oracle/build_ref.py into
// oracle/_ref/libskity_ref.so.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load that library.
//
// What is exercised: skity::Canvas::MakeSoftwareCanvas (src/render/sw/sw_canvas.cc:146),
// SWRaster::RastePath (src/render/sw/sw_raster.cc:731), SWStackBlur
// (src/render/sw/sw_stack_blur.cc:18), exactly as the reference runs them.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <skity/graphic/bitmap.hpp>
#include <skity/render/canvas.hpp>

#include "skity_b200/host/scene_player.hpp"
#include "src/render/sw/sw_raster.hpp"
#include "src/render/sw/sw_stack_blur.hpp"

extern "C" {

// Renders an SKSC scene with the reference software canvas into a premultiplied
// RGBA8 buffer of width*height*4 bytes (row stride width*4, zero-initialised,
// like Bitmap's calloc in src/io/pixmap.cc:77).  `seconds` (optional) receives
// the wall time of the draw loop only.  Returns 0 on success.
int ref_render_scene(const uint8_t* scene, size_t n, uint8_t* out_rgba, double* seconds) {
  if (n < sizeof(skb_scene::Header)) return -1;
  skb_scene::Header h;
  std::memcpy(&h, scene, sizeof(h));
  if (h.magic != skb_scene::kMagic) return -1;
  skity::Bitmap bitmap(h.width, h.height, skity::AlphaType::kPremul_AlphaType);
  if (bitmap.GetPixelAddr() == nullptr) return -6;
  auto canvas = skity::Canvas::MakeSoftwareCanvas(&bitmap);
  if (!canvas) return -7;
  auto t0 = std::chrono::steady_clock::now();
  int rc = skb_scene::Play(scene, n, canvas.get());
  canvas->Flush();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (rc != 0) return rc;
  for (uint32_t y = 0; y < h.height; y++) {
    std::memcpy(out_rgba + static_cast<size_t>(y) * h.width * 4,
                bitmap.GetPixelAddr() + static_cast<size_t>(y) * bitmap.RowBytes(),
                static_cast<size_t>(h.width) * 4);
  }
  return 0;
}

// Runs SWRaster::RastePath on one path (SKSC path record) and returns the span
// list it produced, in emission order.  matrix6 = sx kx tx ky sy ty.
// clip = l t r b.  spans_out receives up to cap spans as int32[4] = x,y,len,cover.
// bounds_out = l t r b of SWRaster::GetBounds().  Returns the span count (may
// exceed cap; only cap are written) or a negative error.
long ref_raster_path(const uint8_t* path_rec, size_t n, const float* matrix6, const float* clip,
                     int32_t* spans_out, long cap, float* bounds_out) {
  skb_scene::Reader r(path_rec, n);
  skity::Path path;
  if (!skb_scene::ReadPath(r, &path)) return -1;
  skity::SWRaster raster;
  raster.RastePath(path, skb_scene::Affine(matrix6),
                   skity::Rect::MakeLTRB(clip[0], clip[1], clip[2], clip[3]));
  auto const& spans = raster.CurrentSpans();
  long cnt = static_cast<long>(spans.size());
  for (long i = 0; i < cnt && i < cap; i++) {
    spans_out[4 * i + 0] = spans[i].x;
    spans_out[4 * i + 1] = spans[i].y;
    spans_out[4 * i + 2] = spans[i].len;
    spans_out[4 * i + 3] = spans[i].cover;
  }
  if (bounds_out) {
    auto b = raster.GetBounds();
    bounds_out[0] = b.Left();
    bounds_out[1] = b.Top();
    bounds_out[2] = b.Right();
    bounds_out[3] = b.Bottom();
  }
  return cnt;
}

// SWStackBlur on a premultiplied RGBA8 buffer (w*h*4, tight rows).
int ref_stack_blur(const uint8_t* src_rgba, uint32_t w, uint32_t h, int radius, uint8_t* dst_rgba) {
  skity::Bitmap src(w, h, skity::AlphaType::kPremul_AlphaType);
  skity::Bitmap dst(w, h, skity::AlphaType::kPremul_AlphaType);
  for (uint32_t y = 0; y < h; y++) {
    std::memcpy(src.GetPixelAddr() + static_cast<size_t>(y) * src.RowBytes(),
                src_rgba + static_cast<size_t>(y) * w * 4, static_cast<size_t>(w) * 4);
  }
  skity::SWStackBlur(&src, &dst, radius).Blur();
  for (uint32_t y = 0; y < h; y++) {
    std::memcpy(dst_rgba + static_cast<size_t>(y) * w * 4,
                dst.GetPixelAddr() + static_cast<size_t>(y) * dst.RowBytes(),
                static_cast<size_t>(w) * 4);
  }
  return 0;
}

const char* ref_version() { return "skity-sw-reference (unmodified sources, glm stand-in)"; }

}  // extern "C"
