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
#include "skity_b200/host/skp_player.hpp"
#include "src/render/hw/coverage/coverage_aa_line_encoder.hpp"
#include "src/render/hw/coverage/coverage_aa_tiler.hpp"
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

// Plays a serialized picture (.skp, module/io) onto the reference software canvas of a width x height bitmap under the
// affine matrix m6 (sx kx tx ky sy ty) — the reference side of tests/golden's skp fixtures.
int ref_render_skp(const uint8_t* skp, size_t n, uint32_t width, uint32_t height, const float* m6, uint8_t* out_rgba, double* seconds) {
  skity::Bitmap bitmap(width, height, skity::AlphaType::kPremul_AlphaType);
  if (bitmap.GetPixelAddr() == nullptr) return -6;
  auto canvas = skity::Canvas::MakeSoftwareCanvas(&bitmap);
  if (!canvas) return -7;
  auto t0 = std::chrono::steady_clock::now();
  int rc = skb_skp::Play(skp, n, m6, canvas.get());
  canvas->Flush();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (rc != 0) return rc;
  for (uint32_t y = 0; y < height; y++) {
    std::memcpy(out_rgba + static_cast<size_t>(y) * width * 4, bitmap.GetPixelAddr() + static_cast<size_t>(y) * bitmap.RowBytes(),
                static_cast<size_t>(width) * 4);
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

// The reference's own CoverageAAPathTiler (src/render/hw/coverage/coverage_aa_tiler.cc:73-94) and line encoder
// (coverage_aa_line_encoder.cc:39-75) on one path (SKSC path record): what its GPU backends upload for a
// coverage-AA draw.  tiles_out[i] = tile_x, tile_y, first line (-1: none), line count, backdrop;
// lines_out[j] = from_x, from_y, to_x, to_y (unsigned 8.8) in the encoded (per-tile) order.  scissor4 may be null.
// Returns the tile count (negative: capacity exceeded); *n_lines_out = encoded lines.
long ref_coverage_aa_tile(const uint8_t* path_rec, size_t n, const float* matrix6, const float* scissor4, int32_t* tiles_out,
                          long tile_cap, uint16_t* lines_out, long line_cap, long* n_lines_out) {
  skb_scene::Reader r(path_rec, n);
  skity::Path path;
  if (!skb_scene::ReadPath(r, &path)) return -2;
  std::vector<skity::CoverageAATile> tiles;
  std::vector<skity::CoverageAATileLine> lines;
  std::vector<uint32_t> counts;
  skity::CoverageAAPathTiler tiler(tiles, lines, counts);
  skity::Rect sc;
  if (scissor4) sc = skity::Rect::MakeLTRB(scissor4[0], scissor4[1], scissor4[2], scissor4[3]);
  skity::CoverageAATiledPath tp = tiler.Tile(path, skb_scene::Affine(matrix6), scissor4 ? &sc : nullptr);
  std::vector<uint32_t> range_counts = counts;   // EncodeCoverageAALines consumes its copy
  skity::CoverageAAEncodedLines enc;
  skity::EncodeCoverageAALines(lines, counts, 16384, enc);
  if (static_cast<long>(tp.tile_count) > tile_cap || static_cast<long>(lines.size()) > line_cap) return -1;
  for (size_t i = 0; i < tp.tile_count; i++) {
    const auto& t = tiles[tp.tile_offset + i];
    int32_t* o = tiles_out + 5 * i;
    o[0] = t.tile_x;
    o[1] = t.tile_y;
    o[2] = t.line_range_id.IsValid() ? static_cast<int32_t>(enc.range_offsets[t.line_range_id.value]) : -1;
    o[3] = t.line_range_id.IsValid() ? static_cast<int32_t>(range_counts[t.line_range_id.value]) : 0;
    o[4] = t.backdrop;
  }
  std::memcpy(lines_out, enc.texture_data.data(), lines.size() * 4 * sizeof(uint16_t));
  *n_lines_out = static_cast<long>(lines.size());
  return static_cast<long>(tp.tile_count);
}

const char* ref_version() { return "skity-sw-reference (unmodified sources, glm stand-in)"; }

}  // extern "C"
