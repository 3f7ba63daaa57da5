// Scene replay: decodes an "SKSC" scene blob and issues the equivalent calls on
// any skity::Canvas.  This is what lets one seeded scene drive both the
// reference software canvas (oracle/_ref) and the CUDA canvas unchanged — the
// same role GoldenTestEnv::RenderToTexture + DisplayList::Draw(canvas) play in
// the reference (test/golden/common/golden_test_env.hpp:45-47,
// src/recorder/display_list.cc:330-343).
//
// Format (little-endian, 4-byte aligned), written by skity_b200/scene.py:
//   header : u32 magic 'SKSC', u32 version, u32 width, u32 height, u32 n_ops, u32 0
//   op     : u32 opcode, u32 payload_bytes, payload
//   path   : u32 fill_type(0 winding,1 even-odd), u32 n_verbs, u32 n_pts, u32 n_weights,
//            u8 verbs[n_verbs] (padded to 4), f32 xy[2*n_pts], f32 w[n_weights]
//            verbs use skity::Path::Verb numbering (include/skity/graphic/path.hpp:45-60)
//   paint  : u32 style, f32 stroke_width, f32 miter, u32 cap, u32 join,
//            f32 fill_rgba[4], f32 stroke_rgba[4], u32 blur_style(0 none,1 normal,2 solid,
//            3 outer,4 inner), f32 blur_radius, u32 shader(0 none,1 linear,2 radial,3 sweep,
//            4 two-point conical: f32 r0, r1 follow the stops)
//            [shader: f32 p[4], u32 tile_mode, u32 n_colors, u32 n_stops, u32 has_local,
//             f32 local[6] (sx kx tx ky sy ty), f32 rgba[4*n_colors], f32 stops[n_stops]]
//            style bit 8 set => extras follow the shader block: u32 blend_mode (skity::BlendMode),
//             u32 image_filter (0 none, 1 ImageFilters::Blur, 2 ImageFilters::DropShadow, 3 ImageFilters::Dilate, 4 Erode: radii in sigma_x/y,
//             5 ImageFilters::MatrixTransform(Translate(dx, dy))),
//             f32 dx, f32 dy, f32 sigma_x, f32 sigma_y, u32 shadow colour (A<<24|R<<16|G<<8|B),
//             u32 colour_filter (0 none, 1 ColorFilters::Blend, 2 Matrix, 3 LinearToSRGBGamma, 4 SRGBToLinearGamma),
//             u32 filter colour, u32 filter blend mode, f32 matrix[20]
#ifndef SKITY_B200_HOST_SCENE_PLAYER_HPP
#define SKITY_B200_HOST_SCENE_PLAYER_HPP

#include <cstdint>
#include <cstring>
#include <memory>
#include <skity/effect/color_filter.hpp>
#include <skity/effect/image_filter.hpp>
#include <skity/effect/mask_filter.hpp>
#include <skity/effect/shader.hpp>
#include <skity/geometry/matrix.hpp>
#include <skity/graphic/bitmap.hpp>
#include <skity/graphic/image.hpp>
#include <skity/graphic/paint.hpp>
#include <skity/graphic/sampling_options.hpp>
#include <skity/graphic/path.hpp>
#include <skity/render/canvas.hpp>
#include <vector>

namespace skb_scene {

enum Op : uint32_t {
  kSave = 1,
  kRestore = 2,
  kTranslate = 3,
  kScale = 4,
  kRotate = 5,
  kConcat = 6,
  kClipRect = 7,
  kClipPath = 8,
  kDrawPath = 9,
  kDrawRect = 10,
  kSaveLayer = 11,  // f32 ltrb[4], paint -> Canvas::SaveLayer(bounds, paint); closed by kRestore
  kDrawImageRect = 12,  // u32 w, h, seed, unpremul (test image), f32 src[4], f32 dst[4], u32 filter, paint
};

constexpr uint32_t kMagic = 0x43534B53u;  // "SKSC"

struct Header {
  uint32_t magic, version, width, height, n_ops, reserved;
};

class Reader {
 public:
  Reader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  bool ok() const { return ok_; }
  uint32_t U32() {
    uint32_t v = 0;
    Get(&v, 4);
    return v;
  }
  float F32() {
    float v = 0;
    Get(&v, 4);
    return v;
  }
  void Get(void* dst, size_t n) {
    if (static_cast<size_t>(end_ - p_) < n) {
      ok_ = false;
      std::memset(dst, 0, n);
      return;
    }
    std::memcpy(dst, p_, n);
    p_ += n;
  }
  void Skip(size_t n) {
    if (static_cast<size_t>(end_ - p_) < n) {
      ok_ = false;
      return;
    }
    p_ += n;
  }
  const uint8_t* Ptr() const { return p_; }

 private:
  const uint8_t* p_;
  const uint8_t* end_;
  bool ok_ = true;
};

inline bool ReadPath(Reader& r, skity::Path* path) {
  uint32_t fill = r.U32(), nv = r.U32(), np = r.U32(), nw = r.U32();
  if (!r.ok()) return false;
  std::vector<uint8_t> verbs((nv + 3) & ~3u);
  r.Get(verbs.data(), verbs.size());
  std::vector<float> pts(2 * static_cast<size_t>(np));
  r.Get(pts.data(), pts.size() * 4);
  std::vector<float> ws(nw);
  r.Get(ws.data(), ws.size() * 4);
  if (!r.ok()) return false;
  size_t pi = 0, wi = 0;
  for (uint32_t i = 0; i < nv; i++) {
    switch (verbs[i]) {
      case 0:
        if (pi + 1 > np) return false;
        path->MoveTo(pts[2 * pi], pts[2 * pi + 1]);
        pi += 1;
        break;
      case 1:
        if (pi + 1 > np) return false;
        path->LineTo(pts[2 * pi], pts[2 * pi + 1]);
        pi += 1;
        break;
      case 2:
        if (pi + 2 > np) return false;
        path->QuadTo(pts[2 * pi], pts[2 * pi + 1], pts[2 * pi + 2], pts[2 * pi + 3]);
        pi += 2;
        break;
      case 3:
        if (pi + 2 > np || wi + 1 > nw) return false;
        path->ConicTo(pts[2 * pi], pts[2 * pi + 1], pts[2 * pi + 2], pts[2 * pi + 3], ws[wi]);
        pi += 2;
        wi += 1;
        break;
      case 4:
        if (pi + 3 > np) return false;
        path->CubicTo(pts[2 * pi], pts[2 * pi + 1], pts[2 * pi + 2], pts[2 * pi + 3],
                      pts[2 * pi + 4], pts[2 * pi + 5]);
        pi += 3;
        break;
      case 5:
        path->Close();
        break;
      default:
        return false;
    }
  }
  path->SetFillType(fill == 1 ? skity::Path::PathFillType::kEvenOdd
                              : skity::Path::PathFillType::kWinding);
  return true;
}

// A deterministic test image: a smooth colour field with hard-edged translucent blocks, so that nearest and bilinear
// sampling, tiling and both alpha types all show.  Pixels are valid premultiplied colours when `unpremul` is 0.
inline std::shared_ptr<skity::Image> MakeTestImage(uint32_t w, uint32_t h, uint32_t seed, uint32_t unpremul) {
  skity::Bitmap bm(w, h, unpremul ? skity::AlphaType::kUnpremul_AlphaType : skity::AlphaType::kPremul_AlphaType);
  uint32_t state = seed * 2654435761u + 12345u;
  for (uint32_t y = 0; y < h; y++) {
    for (uint32_t x = 0; x < w; x++) {
      uint32_t r = (x * 255u) / (w > 1 ? w - 1 : 1), g = (y * 255u) / (h > 1 ? h - 1 : 1);
      uint32_t b = ((x ^ y) * 37u + seed * 11u) & 0xFF;
      uint32_t a = 255;
      if (((x / 5) + (y / 3) + seed) % 4 == 0) {
        state = state * 1664525u + 1013904223u;
        a = (state >> 24) & 0xFF;
      }
      if (!unpremul) {
        r = r * a / 255u;
        g = g * a / 255u;
        b = b * a / 255u;
      }
      bm.SetPixel(x, y, skity::ColorSetARGB(static_cast<uint8_t>(a), static_cast<uint8_t>(r), static_cast<uint8_t>(g),
                                            static_cast<uint8_t>(b)));
    }
  }
  return skity::Image::MakeImage(bm.GetPixmap());
}

inline skity::Matrix Affine(const float m[6]) {
  // m = sx kx tx ky sy ty (row-major 2x3)
  return skity::Matrix(m[0], m[1], m[2], m[3], m[4], m[5], 0.f, 0.f, 1.f);
}

inline bool ReadPaint(Reader& r, skity::Paint* paint) {
  uint32_t style = r.U32();
  float sw = r.F32(), miter = r.F32();
  uint32_t cap = r.U32(), join = r.U32();
  float fc[4], sc[4];
  r.Get(fc, 16);
  r.Get(sc, 16);
  uint32_t blur_style = r.U32();
  float blur_radius = r.F32();
  uint32_t shader = r.U32();
  if (!r.ok()) return false;
  const bool extras = (style & 0x100u) != 0;
  style &= 0xFFu;
  paint->SetStyle(static_cast<skity::Paint::Style>(style));
  paint->SetStrokeWidth(sw);
  paint->SetStrokeMiter(miter);
  paint->SetStrokeCap(static_cast<skity::Paint::Cap>(cap));
  paint->SetStrokeJoin(static_cast<skity::Paint::Join>(join));
  paint->SetFillColor(fc[0], fc[1], fc[2], fc[3]);
  paint->SetStrokeColor(sc[0], sc[1], sc[2], sc[3]);
  if (blur_style != 0) {
    paint->SetMaskFilter(
        skity::MaskFilter::MakeBlur(static_cast<skity::BlurStyle>(blur_style), blur_radius));
  }
  if (shader != 0) {
    float p[4];
    r.Get(p, 16);
    uint32_t tile = r.U32(), nc = r.U32(), ns = r.U32(), has_local = r.U32();
    float local[6];
    r.Get(local, 24);
    if (!r.ok() || nc < 2 || nc > 4096 || (ns != 0 && ns != nc)) return false;
    std::vector<skity::Vec4> colors(nc);
    for (uint32_t i = 0; i < nc; i++) {
      float c[4];
      r.Get(c, 16);
      colors[i] = skity::Vec4{c[0], c[1], c[2], c[3]};
    }
    std::vector<float> stops(ns);
    r.Get(stops.data(), ns * 4);
    if (!r.ok()) return false;
    const float* pos = ns ? stops.data() : nullptr;
    auto tm = static_cast<skity::TileMode>(tile);
    std::shared_ptr<skity::Shader> sh;
    if (shader == 1) {
      skity::Point pts[2] = {skity::Point{p[0], p[1], 0.f, 1.f}, skity::Point{p[2], p[3], 0.f, 1.f}};
      sh = skity::Shader::MakeLinear(pts, colors.data(), pos, static_cast<int>(nc), tm);
    } else if (shader == 2) {
      sh = skity::Shader::MakeRadial(skity::Point{p[0], p[1], 0.f, 1.f}, p[2], colors.data(), pos,
                                     static_cast<int>(nc), tm);
    } else if (shader == 3) {
      sh = skity::Shader::MakeSweep(p[0], p[1], p[2], p[3], colors.data(), pos,
                                    static_cast<int>(nc), tm);
    } else if (shader == 4) {  // two-point conical: p = start.xy, end.xy; radii follow the stops
      float rr[2];
      r.Get(rr, 8);
      if (!r.ok()) return false;
      sh = skity::Shader::MakeTwoPointConical(skity::Point{p[0], p[1], 0.f, 1.f}, rr[0], skity::Point{p[2], p[3], 0.f, 1.f},
                                              rr[1], colors.data(), pos, static_cast<int>(nc), tm);
    } else if (shader == 5) {  // image shader: p = width, height, seed, unpremul; y tile mode and filter follow the stops
      uint32_t ymode = r.U32(), filter = r.U32();
      if (!r.ok() || ymode > 3 || filter > 2 || p[0] < 1 || p[1] < 1 || p[0] > 4096 || p[1] > 4096) return false;
      skity::SamplingOptions so;
      so.filter = filter == 0 ? skity::FilterMode::kNearest : skity::FilterMode::kLinear;
      if (filter == 2) so.cubic = skity::CubicResampler{1 / 3.0f, 1 / 3.0f};
      sh = skity::Shader::MakeShader(MakeTestImage(static_cast<uint32_t>(p[0]), static_cast<uint32_t>(p[1]),
                                                   static_cast<uint32_t>(p[2]), static_cast<uint32_t>(p[3])),
                                     so, tm, static_cast<skity::TileMode>(ymode),
                                     has_local ? Affine(local) : skity::Matrix());
      has_local = 0;
    } else {
      return false;
    }
    if (sh && has_local) sh->SetLocalMatrix(Affine(local));
    paint->SetShader(sh);
  }
  if (extras) {
    uint32_t blend = r.U32();
    uint32_t filter = r.U32();
    float f[4];
    r.Get(f, 16);
    uint32_t shadow = r.U32();
    if (!r.ok() || blend > static_cast<uint32_t>(skity::BlendMode::kLastMode) || filter > 5) return false;
    paint->SetBlendMode(static_cast<skity::BlendMode>(blend));
    if (filter == 1) paint->SetImageFilter(skity::ImageFilters::Blur(f[2], f[3]));
    if (filter == 2) paint->SetImageFilter(skity::ImageFilters::DropShadow(f[0], f[1], f[2], f[3], shadow, nullptr));
    if (filter == 3) paint->SetImageFilter(skity::ImageFilters::Dilate(f[2], f[3]));
    if (filter == 4) paint->SetImageFilter(skity::ImageFilters::Erode(f[2], f[3]));
    if (filter == 5) paint->SetImageFilter(skity::ImageFilters::MatrixTransform(skity::Matrix::Translate(f[0], f[1])));
    uint32_t cf = r.U32(), cf_color = r.U32(), cf_mode = r.U32();
    float cm[20];
    r.Get(cm, 80);
    if (!r.ok() || cf > 4 || cf_mode > static_cast<uint32_t>(skity::BlendMode::kLastMode)) return false;
    if (cf == 1) paint->SetColorFilter(skity::ColorFilters::Blend(cf_color, static_cast<skity::BlendMode>(cf_mode)));
    if (cf == 2) paint->SetColorFilter(skity::ColorFilters::Matrix(cm));
    if (cf == 3) paint->SetColorFilter(skity::ColorFilters::LinearToSRGBGamma());
    if (cf == 4) paint->SetColorFilter(skity::ColorFilters::SRGBToLinearGamma());
  }
  return true;
}

// Returns 0 on success, negative on a malformed blob.
inline int Play(const uint8_t* data, size_t n, skity::Canvas* canvas) {
  Reader r(data, n);
  Header h;
  r.Get(&h, sizeof(h));
  if (!r.ok() || h.magic != kMagic || h.version != 1) return -1;
  for (uint32_t i = 0; i < h.n_ops; i++) {
    uint32_t op = r.U32();
    uint32_t bytes = r.U32();
    if (!r.ok()) return -2;
    const uint8_t* start = r.Ptr();
    switch (op) {
      case kSave:
        canvas->Save();
        break;
      case kRestore:
        canvas->Restore();
        break;
      case kTranslate: {
        float a = r.F32(), b = r.F32();
        canvas->Translate(a, b);
      } break;
      case kScale: {
        float a = r.F32(), b = r.F32();
        canvas->Scale(a, b);
      } break;
      case kRotate:
        canvas->Rotate(r.F32());
        break;
      case kConcat: {
        float m[6];
        r.Get(m, 24);
        canvas->Concat(Affine(m));
      } break;
      case kClipRect: {
        float q[4];
        r.Get(q, 16);
        uint32_t cop = r.U32();
        canvas->ClipRect(skity::Rect::MakeLTRB(q[0], q[1], q[2], q[3]),
                         cop ? skity::Canvas::ClipOp::kIntersect : skity::Canvas::ClipOp::kDifference);
      } break;
      case kClipPath: {
        skity::Path path;
        if (!ReadPath(r, &path)) return -3;
        uint32_t cop = r.U32();
        canvas->ClipPath(path, cop ? skity::Canvas::ClipOp::kIntersect
                                   : skity::Canvas::ClipOp::kDifference);
      } break;
      case kDrawPath: {
        skity::Path path;
        skity::Paint paint;
        if (!ReadPath(r, &path) || !ReadPaint(r, &paint)) return -3;
        canvas->DrawPath(path, paint);
      } break;
      case kDrawRect: {
        float q[4];
        r.Get(q, 16);
        skity::Paint paint;
        if (!ReadPaint(r, &paint)) return -3;
        canvas->DrawRect(skity::Rect::MakeLTRB(q[0], q[1], q[2], q[3]), paint);
      } break;
      case kDrawImageRect: {
        uint32_t iw = r.U32(), ih = r.U32(), seed = r.U32(), unpremul = r.U32();
        float sq[4], dq[4];
        r.Get(sq, 16);
        r.Get(dq, 16);
        uint32_t filter = r.U32();
        skity::Paint paint;
        if (!ReadPaint(r, &paint) || iw < 1 || ih < 1 || iw > 4096 || ih > 4096 || filter > 2) return -3;
        skity::SamplingOptions so;
        so.filter = filter == 0 ? skity::FilterMode::kNearest : skity::FilterMode::kLinear;
        if (filter == 2) so.cubic = skity::CubicResampler{1 / 3.0f, 1 / 3.0f};
        canvas->DrawImageRect(MakeTestImage(iw, ih, seed, unpremul), skity::Rect::MakeLTRB(sq[0], sq[1], sq[2], sq[3]),
                              skity::Rect::MakeLTRB(dq[0], dq[1], dq[2], dq[3]), so, &paint);
      } break;
      case kSaveLayer: {
        float q[4];
        r.Get(q, 16);
        skity::Paint paint;
        if (!ReadPaint(r, &paint)) return -3;
        canvas->SaveLayer(skity::Rect::MakeLTRB(q[0], q[1], q[2], q[3]), paint);
      } break;
      default:
        return -4;
    }
    if (!r.ok() || static_cast<size_t>(r.Ptr() - start) != bytes) return -5;
  }
  return 0;
}

}  // namespace skb_scene

#endif  // SKITY_B200_HOST_SCENE_PLAYER_HPP
