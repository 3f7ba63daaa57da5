/*
 * skb_dl.h — the flat, device-ready display list ("SKDL") consumed by the CUDA
 * backend through skb_frame_encode() (include/skb.h).
 *
 * It is what the host-side canvas (skity_b200/host/cuda_canvas.cc) produces from
 * skity::Canvas calls, the way the reference's RecordingCanvas turns the same
 * calls into DisplayList ops (src/recorder/recorded_op.hpp:19-56).  Everything
 * in it is already lowered to what the reference's software canvas hands to
 * SWRaster::RastePath / SWSpanBrush (src/render/sw/sw_canvas.cc:357-411,727-826):
 * fills only (strokes are expanded on the host by the reference's own Stroke,
 * exactly as both existing backends do — sw_canvas.cc:388-401, hw_canvas.cc:476-485),
 * the CTM, the scan clip rectangle, the clip-stack state and the brush.
 *
 * Plain C, little-endian, every section 16-byte aligned.  No pointers.
 */
#ifndef SKB_DL_H
#define SKB_DL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_DL_MAGIC 0x4C444B53u /* "SKDL" */
#define SKB_DL_VERSION 1u

typedef struct skb_dl_header {
  uint32_t magic;
  uint32_t version;
  uint32_t total_bytes;
  uint32_t flags;
  uint32_t n_surfaces; /* surface 0 is the canvas; others are offscreen temporaries */
  uint32_t n_ops;
  uint32_t n_paths;
  uint32_t n_segs;
  uint32_t n_paints;
  uint32_t n_stop_floats; /* floats in the gradient colour/stop pool */
  uint32_t n_clip_states; /* clip state ids are 1..n_clip_states; 0 = unclipped */
  uint32_t reserved0;
  uint32_t off_surfaces; /* byte offsets from the start of the blob */
  uint32_t off_ops;
  uint32_t off_paths;
  uint32_t off_segs;
  uint32_t off_paints;
  uint32_t off_stops;
  uint32_t reserved1[2];
} skb_dl_header; /* 80 bytes */

/* flags bit 0: the surface is a CANVAS of a batch (a final image the caller reads back with
 * skb_frame_read_surface), not a blur temporary.  Canvases are composited after the blur stage,
 * like surface 0.  A batch of independent canvases is one display list whose ops target surfaces
 * 1..N: all canvases share every launch, which is what makes many small canvases efficient. */
#define SKB_SURFACE_CANVAS 1u

/* skb_dl_surface.flags */
#define SKB_SURFACE_IMAGE 2u  /* an application image (Image::MakeImage of a Pixmap): its pixels travel in the display
                                 list, `reserved` = byte offset from the start of the list of width*height RGBA8 pixels
                                 (the Colors Bitmap::GetPixel returns, src/graphic/bitmap.cc:25-50, as R,G,B,A bytes) */

typedef struct skb_dl_surface {
  uint32_t width;
  uint32_t height;
  uint32_t flags;
  uint32_t reserved;
} skb_dl_surface;

enum skb_dl_op_kind {
  SKB_OP_FILL = 1, /* SWRaster::RastePath + SWSpanBrush::Brush of one path */
  SKB_OP_CLIP = 2, /* SWCanvas::OnClipPath: rasterise and combine into a new clip state */
  SKB_OP_BLUR = 3  /* SWStackBlur: surface aux -> surface `surface`, radius in clip_bounds[0]; fill_type = what is
                      then done to the blurred pixels with the unblurred ones at hand: 0 nothing (BlurStyle::kNormal,
                      ImageFilters::Blur), 2 kSolid, 3 kOuter, 4 kInner (src/effect/mask_filter.cc:64-100),
                      5 drop shadow (src/effect/image_filter.cc:222-233) with the colour in `paint`;
                      6 / 7: no blur at all but ImageFilters::Dilate / Erode (MorphologyImageFilter::OnFilter,
                      image_filter.cc:294-385): clip_bounds[0], [1] = the filter's radius_x, radius_y */
};

typedef struct skb_dl_op {
  uint32_t kind;
  uint32_t surface;   /* FILL/CLIP: target surface; BLUR: destination surface */
  uint32_t path;      /* FILL/CLIP */
  uint32_t paint;     /* FILL; BLUR style 5: the shadow's skity::Color (A<<24|R<<16|G<<8|B, unpremultiplied) */
  uint32_t clip_in;   /* FILL: clip state applied; CLIP: state being refined (0 = none) */
  uint32_t clip_out;  /* CLIP: id of the state this op defines */
  uint32_t fill_type; /* 0 nonzero winding, 1 even-odd (Path::PathFillType); BLUR: style, see SKB_OP_BLUR */
  uint32_t aux;       /* CLIP: Canvas::ClipOp (0 difference, 1 intersect); BLUR: source surface */
  float ctm[6];       /* sx kx tx ky sy ty — SWCanvas::CurrentTransform() */
  float clip_bounds[4]; /* l t r b — SWCanvas::GetScanClipBounds(); BLUR: [0] = integer radius */
} skb_dl_op; /* 72 bytes */

typedef struct skb_dl_path {
  uint32_t seg_off;
  uint32_t n_segs;
  uint32_t reserved[2];
} skb_dl_path;

/* One entry per segment of the lowered path, i.e. of the Path that
 * Stroke::QuadPath(src, keep curves) followed by PathEdgeIter would walk
 * (src/geometry/stroke.cc:914-962, src/graphic/path_priv.hpp:75-167): explicit
 * start point, implicit closing lines materialised. */
enum skb_dl_seg_type {
  SKB_SEG_POINT = 0, /* contributes to the path bounds only (a lone MoveTo that stays in the path) */
  SKB_SEG_LINE = 1,
  SKB_SEG_QUAD = 2,
  SKB_SEG_CONIC = 3,
  SKB_SEG_CUBIC = 4,
  SKB_SEG_CLOSE = 5 /* auto-close line p0 -> p1; its end points are not new bounds points */
};
/* flag: the segment STARTS at the COMPUTED end of the preceding cubic
 * (CubicCoeff::EvalAt(1)) instead of start[] — Cubic::ToQuads chains its quads
 * through the destination path (src/geometry/cubic.cc:29-52), so whatever
 * follows a cubic begins where the last emitted quad ended. */
#define SKB_SEG_P0_FROM_PREV_CUBIC 0x100u
#define SKB_SEG_TYPE_MASK 0xFFu

typedef struct skb_dl_seg {
  uint32_t type_flags;
  float w;        /* conic weight */
  float p[8];     /* x0 y0 x1 y1 x2 y2 x3 y3 in path (pre-CTM) space: the points Path::Iter hands to
                     Stroke::QuadPath, used for the curve maths (p0 is the SOURCE path's previous point) */
  float start[2]; /* where the segment begins in the lowered path (its last point so far) */
} skb_dl_seg;     /* 48 bytes */

enum skb_dl_paint_type {
  SKB_PAINT_SOLID = 0,
  SKB_PAINT_LINEAR = 1,
  SKB_PAINT_RADIAL = 2,
  SKB_PAINT_SWEEP = 3,
  SKB_PAINT_IMAGE = 4, /* PixmapBrush, nearest, decal/decal (the blur composite) */
  SKB_PAINT_CONICAL = 5 /* two-point conical gradient: m = device_to_local; the constants
                           ConicalGradientColorBrush::OnPreBrush derives (sw_span_brush.cc:405-444) follow the
                           stops in the float pool, 16 floats: kind (0 transparent, 1 concentric, 2 equal radii,
                           3 general, 5 general with the circles swapped), c0.x, c0.y, scale, scale_sign, bias,
                           r0/|c1-c0|, transform sx kx tx ky sy ty, r1, r1^2, f */
};

/* Colour-filter block in the float pool (raw 32-bit words), applied to the coverage-scaled source colour before
 * the blend (SWSpanBrush::BrushH, sw_span_brush.cc:108-133; src/effect/color_filter.cc:123-197):
 *   word 0 type: SKB_CF_BLEND   word 1 = skity::BlendMode, word 2 = the filter's premultiplied colour A<<24|R<<16|G<<8|B
 *                SKB_CF_MATRIX  words 4-13 = the 4x5 matrix as int16 (row-major, MatrixColorFilter::matrix_i16_)
 *                SKB_CF_TABLE   words 4-67 = 256 bytes: per-channel look-up of the unpremultiplied colour
 *                               (SRGBGammaColorFilter, either direction) */
#define SKB_CF_BLEND 1u
#define SKB_CF_MATRIX 2u
#define SKB_CF_TABLE 3u
#define SKB_PAINT_HAS_STOPS(p) ((p).has_stops & 1u)
#define SKB_PAINT_CF_OFFSET(p) ((p).has_stops >> 8) /* 0 = none, else 1 + word offset */

/* IMAGE paints, ORed into tile_mode (whose low nibble is the x tile mode):
 *   UNPREMUL  the sampled surface holds unpremultiplied pixels (PixmapBrush premultiplies after sampling,
 *             sw_span_brush.cc:573-576)
 *   LINEAR    FilterMode::kLinear (BitmapSampler::SampleUnitLinear, bitmap_sampler.cc:44-83); cubic resampling
 *             falls back to it in the reference (:96-99)
 *   YMODE     bits 4-7 hold the y tile mode (otherwise it equals the x mode) */
#define SKB_PAINT_IMAGE_UNPREMUL 0x100u
#define SKB_PAINT_IMAGE_LINEAR 0x200u
#define SKB_PAINT_IMAGE_YMODE 0x400u

typedef struct skb_dl_paint {
  uint32_t type;
  uint32_t tile_mode; /* skity::TileMode: 0 clamp 1 repeat 2 mirror 3 decal (low byte) */
  float color[4];     /* SOLID: unpremultiplied r g b a as Paint holds them (Color4f) */
  float m[6];         /* gradients: PointsToUnit * device_to_local; IMAGE: Scale(1/w,1/h) * inv(local) * inv(CTM)
                         — sx kx tx ky sy ty, as GenerateBrush builds it (sw_canvas.cc:727-795) */
  uint32_t stop_off;  /* float offset into the stop pool: n_colors*4 colour floats then n_colors stops */
  uint32_t n_colors;
  uint32_t has_stops; /* bit 0: explicit stops (else implicit i/(n-1)); bits 8-31: 1 + word offset in the float pool
                         of a colour-filter block (0 = no filter), see SKB_CF_* */
  float bias;         /* SWEEP: info.radius[0] */
  float scale;        /* SWEEP: info.radius[1] */
  uint32_t image_surface; /* IMAGE: source surface id */
  uint32_t global_alpha;  /* uint8(255*alpha), ANDed with coverage (sw_span_brush.cc:101) */
  uint32_t blend;         /* 0 = kSrcOver (the default); otherwise skity::BlendMode value + 1.  Modes
                             SWRenderTarget does not implement are sent as kSrcOver, its own fall-back
                             (blend_mode.cc:129-133) */
} skb_dl_paint; /* 80 bytes */

#ifdef __cplusplus
}
#endif

#endif /* SKB_DL_H */
