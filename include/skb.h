/*
 * skb.h — C ABI of the B200 rasterisation backend (libskb.so).
 *
 * This is the boundary the CUDA side exports and the only thing the host-side
 * skity plug-in (skity_b200/host/) binds.  Conventions follow the reference's
 * own C API (module/capi): opaque handles `typedef struct x_s* x`
 * (include/skity_c/skity_base.h:34), every fallible call returns a result enum
 * that is 0 on success and negative on failure (:47-56), out-parameters last,
 * create/destroy pairs, no exceptions cross the boundary.
 *
 * What each entry point replaces in the reference (paths relative to its root):
 *   skb_device_create / destroy     GLContextCreate(void*) and the GPUContext lifetime
 *                                   (include/skity/gpu/gpu_context_gl.hpp:115, src/gpu/gl/gpu_context_impl_gl.cc:74-82)
 *   skb_surface_create / destroy    GPUContext::CreateSurface(GPUSurfaceDescriptor*) (include/skity/gpu/gpu_context.hpp:73-90);
 *                                   on the CPU path: Bitmap(w,h,kPremul) + Canvas::MakeSoftwareCanvas (src/render/sw/sw_canvas.cc:146-156)
 *   skb_frame_begin                 GPUSurface::LockCanvas(bool clear) (include/skity/gpu/gpu_surface.hpp:96-104)
 *   skb_frame_encode                the draw calls of one frame: what SWCanvas::OnDrawPath/OnClipPath/HandleFilter
 *                                   hand to SWRaster::RastePath + SWSpanBrush::Brush + SWStackBlur
 *                                   (src/render/sw/sw_canvas.cc:315-411,797-826), as a flat list (include/skb_dl.h)
 *   skb_frame_flush                 Canvas::Flush / GPUSurface::Flush (include/skity/gpu/gpu_surface.hpp:106-111)
 *   skb_surface_read_pixels         GPUSurface::ReadPixels(const Rect&) (include/skity/gpu/gpu_surface.hpp:113-128)
 *   skb_surface_device_ptr          (no equivalent) device address of the band, for the NCCL gather
 *
 * Pixels are premultiplied RGBA8 (bytes R,G,B,A), the configuration the
 * reference's Bitmap(AlphaType::kPremul_AlphaType, ColorType::kRGBA) has.
 * Threading: like GPUContext (include/skity/gpu/gpu_context.hpp:65,101) a surface is driven by one thread at a
 * time; different surfaces of a device may be driven by different threads (each has its own stream and arenas,
 * the last-error string is per thread).
 */
#ifndef SKB_H
#define SKB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SKB_API
#else
#define SKB_API __attribute__((visibility("default")))
#endif

typedef struct skb_device_s* skb_device;
typedef struct skb_surface_s* skb_surface;

typedef enum skb_result {
  SKB_SUCCESS = 0,
  SKB_ERROR_INVALID_ARGUMENT = -1,
  SKB_ERROR_CUDA = -2,          /* a CUDA runtime call or kernel failed; see skb_get_last_error_string() */
  SKB_ERROR_OUT_OF_MEMORY = -3,
  SKB_ERROR_UNSUPPORTED = -4,   /* the display list uses something outside the backend's scope */
  SKB_ERROR_BAD_DISPLAY_LIST = -5,
  SKB_ERROR_NO_DEVICE = -6      /* no usable CUDA device: there is no CPU fallback */
} skb_result;

/* Per-frame counters and device timings (CUDA events on the surface's stream). */
typedef struct skb_frame_stats {
  uint32_t n_ops, n_segs, n_prims, n_edges_slots;
  uint64_t n_rows, n_records, n_items, n_items_nonempty, n_cmds, n_tiles;
  uint64_t pool_capacity;
  uint32_t n_launches;   /* kernels launched by the last flush */
  uint32_t n_retries;    /* record-pool overflows that forced a re-run */
  float ms_total;        /* first launch to last launch, device time */
  float ms_stage[8];     /* 0 flatten, 1 setup+scans, 2 walk, 3 coverage, 4 bin, 5 fine, 6 blur, 7 clip */
  uint64_t bytes_fine;   /* algorithmic bytes of the fine pass (pixels + commands + masks) */
  uint64_t bytes_cover;  /* algorithmic bytes of the coverage pass (records in, masks out) */
  uint64_t bytes_walk;   /* algorithmic bytes of the sweep (edge slots in, records + row table out) */
  uint64_t bytes_blur;   /* algorithmic bytes of the blur passes: 2 passes x (4 B read + 4 B write) per temp pixel */
  uint32_t n_rw_retried;     /* paths whose row-parallel sweep was retried (tables not self-consistent at first) */
  uint32_t n_rw_sequential;  /* paths swept by the sequential walker */
  uint32_t n_area_lines;       /* coverage mode AREA: lines the unclipped fills flattened to */
  uint32_t n_area_tile_lines;  /* ... and the (line, tile) pairs they were binned to */
  uint64_t bytes_area;         /* algorithmic bytes of the AREA coverage pass (binned lines in, masks out) */
} skb_frame_stats;

SKB_API skb_result skb_device_create(int ordinal, skb_device* out_device);
SKB_API void skb_device_destroy(skb_device device);
SKB_API skb_result skb_device_sm_count(skb_device device, int* out_sm_count);

SKB_API skb_result skb_surface_create(skb_device device, uint32_t width, uint32_t height, skb_surface* out_surface);
SKB_API void skb_surface_destroy(skb_surface surface);

/* Restrict rendering to the tile band [y0, y1) of the canvas (rows rounded outwards to 16).
 * Used to split one canvas across GPUs; y1 = 0 resets to the whole surface. */
SKB_API skb_result skb_surface_set_band(skb_surface surface, uint32_t y0, uint32_t y1);

/* How float coordinates become 16.16 fixed point.  The reference converts through 26.6 with `x << 10` in int32
 * (SWFDot6ToFixed, src/render/sw/sw_subpixel.hpp:43, used by SWEdge::SetLine / SWQuadEdge::SetQuad,
 * src/render/sw/sw_edge.cc:22-29,124-130,207-215): at 8192 px the shift overflows and the coordinate wraps, so
 * on canvases larger than 8192 px geometry beyond that line lands elsewhere (the reference's own numeric range).
 *   SKB_COORD_REFERENCE  the reference's arithmetic, wrap included (bit-exact with it for every input)
 *   SKB_COORD_WIDE       the same conversion without the overflow: identical results wherever the reference does
 *                        not wrap, coordinates valid up to +-32767 px
 *   SKB_COORD_AUTO       (default) WIDE for surfaces wider or taller than 8192 px, REFERENCE otherwise */
#define SKB_COORD_AUTO 0
#define SKB_COORD_REFERENCE 1
#define SKB_COORD_WIDE 2
SKB_API skb_result skb_surface_set_coord_mode(skb_surface surface, int mode);

/* Which form of stage 3 (the reference's active-edge sweep, src/render/sw/sw_raster.cc:546-677) runs:
 *   0  one thread per path (default);
 *   1  row-parallel: one thread per (path, pixel row), chords chained through per-path band tables, bit-identical
 *      trapezoid records (skity_b200/csrc/skb_rowwalk.cuh); paths it cannot settle are swept as in mode 0. */
SKB_API skb_result skb_surface_set_walk_mode(skb_surface surface, int mode);

/* How the coverage of UNCLIPPED fills is computed (clip paths and clipped draws always take the exact route):
 *   SKB_COVERAGE_EXACT  (default) the software backend's analytic-AA scan converter, reproduced bit for bit
 *                       (src/render/sw/sw_raster.cc): flatten to 16.16 edges, sweep, trapezoid rows -> A8 masks;
 *   SKB_COVERAGE_AREA   the algorithm of the reference's GPU coverage-AA path (src/render/hw/coverage/
 *                       coverage_aa_tiler.cc, wgsl_coverage_aa_common.hpp): curves flattened by Wang's formula, lines
 *                       binned to 16x16 tiles with backdrop deltas, per-tile signed-area accumulation, backdrop prefix
 *                       sums — fully parallel.  Its A8 coverage equals that algorithm's (bit-exact against the oracle's
 *                       restatement and the reference's exact-match golden canonical_edges_exact.png); it is NOT the
 *                       software backend's coverage: edge pixels differ (tests/ and DESIGN.md give the histogram). */
#define SKB_COVERAGE_EXACT 0
#define SKB_COVERAGE_AREA 1
SKB_API skb_result skb_surface_set_coverage_mode(skb_surface surface, int mode);

/* clear != 0 zeroes the surface (transparent black), like LockCanvas(true). */
SKB_API skb_result skb_frame_begin(skb_surface surface, int clear);
/* Copies the display list to the device (host -> device, asynchronous on the surface's stream
 * when `display_list` is pinned) and validates it.  One display list per frame. */
SKB_API skb_result skb_frame_encode(skb_surface surface, const void* display_list, size_t bytes);
/* The validation skb_frame_encode performs, on its own (no device needed): SKB_SUCCESS, SKB_ERROR_BAD_DISPLAY_LIST or
 * SKB_ERROR_UNSUPPORTED, with the reason in skb_get_last_error_string().  A list that passes is safe to hand to the
 * kernels: every index is in range, sections are in order, every fill / clip op owns the next path and the paths tile
 * the segment table (include/skb_dl.h). */
SKB_API skb_result skb_display_list_validate(const void* display_list, size_t bytes);
/* Host-side band partition of a display list (no device needed): writes to `out` the part of `display_list` that can
 * reach rows [row0, row1) of the canvas — the list a device rendering that band of the canvas (skb_surface_set_band)
 * needs.  Only fills of the canvas whose path cannot reach the rows are dropped, by the same conservative test the
 * device applies before flattening with a wider margin, so the band's pixels are identical to those rendered from the
 * whole list; clip paths, blurs and draws into other surfaces all stay.  `out` null: only `*out_bytes` (the size
 * needed) is set.  This is the multi-GPU counterpart of the reference handing one display list to one canvas
 * (src/recorder/display_list.hpp:41-64, DisplayList::Draw): N devices, N culled lists, instead of N copies of it. */
SKB_API skb_result skb_display_list_cull_rows(const void* display_list, size_t bytes, int32_t row0, int32_t row1, void* out,
                                              size_t out_capacity, size_t* out_bytes);
/* Launches every stage for the encoded frame.  Asynchronous unless the record pool overflows. */
SKB_API skb_result skb_frame_flush(skb_surface surface);
/* Blocks until the surface's stream is idle; returns the first asynchronous error. */
SKB_API skb_result skb_surface_sync(skb_surface surface);

/* Page-locked host memory: display lists encoded into it upload asynchronously at PCIe speed, read-backs into it need no
 * staging.  (The plug-in keeps its per-frame display list in such a buffer.) */
SKB_API skb_result skb_host_alloc(size_t bytes, void** out_ptr);
SKB_API void skb_host_free(void* ptr);

/* Copies the rectangle to host memory (rows `stride` bytes apart).  Synchronises.  A large read into pageable memory
 * (a fresh Pixmap) is staged through two page-locked buffers of the surface, chunk by chunk, with the host copy done by
 * several threads. */
SKB_API skb_result skb_surface_read_pixels(skb_surface surface, uint32_t x, uint32_t y, uint32_t width,
                                           uint32_t height, void* dst, size_t stride);
/* Same copy, enqueued on the surface's stream without waiting: `dst` (pinned host memory for a truly
 * asynchronous copy) is valid after the next skb_surface_sync().  Lets an application keep two
 * surfaces in flight so that the read-back of one frame overlaps the rendering of the next. */
SKB_API skb_result skb_surface_read_pixels_async(skb_surface surface, uint32_t x, uint32_t y, uint32_t width,
                                                 uint32_t height, void* dst, size_t stride);
/* Uploads pixels into the surface (the LockCanvas(false) case: drawing over existing content). */
SKB_API skb_result skb_surface_write_pixels(skb_surface surface, uint32_t x, uint32_t y, uint32_t width,
                                            uint32_t height, const void* src, size_t stride);
/* Reads back surface `index` (> 0) of the last flushed display list — a canvas of a batch
 * (SKB_SURFACE_CANVAS in include/skb_dl.h).  Valid until the next skb_frame_flush. */
SKB_API skb_result skb_frame_read_surface(skb_surface surface, uint32_t index, void* dst, size_t stride);
SKB_API skb_result skb_surface_device_ptr(skb_surface surface, void** out_ptr, size_t* out_pitch_bytes);
/* One canvas rendered by several GPUs (bands, skb_surface_set_band) with the gather fused into the fine pass: the
 * gathering process exports its canvas as a 64-byte CUDA IPC handle; every other process (one per GPU, same
 * surface size) opens it, after which its fine pass stores the finished pixels of its band straight into the
 * gathering GPU's canvas over NVLink (peer memory) instead of its own — no copy, no collective, only a barrier
 * before the read-back.  NULL restores local stores.
 * Preconditions (the caller's, not checked): the fine pass blends against the LOCAL canvas and stores the result to the
 * remote one, and skips tiles nothing was drawn into unless they lie in its band — so every process must start the
 * frame from the same content (skb_frame_begin(clear = 1) on all of them, or identical skb_surface_write_pixels), and a
 * barrier must separate the gathering process's clear from the other processes' skb_frame_flush (skity_b200/multigpu.py:
 * fuse_gather_into_fine_pass, tests/multigpu_check.py). */
SKB_API skb_result skb_surface_export_canvas(skb_surface surface, void* out_handle64);
SKB_API skb_result skb_surface_set_remote_canvas(skb_surface surface, const void* handle64);
SKB_API skb_result skb_surface_stream(skb_surface surface, void** out_cuda_stream);

SKB_API skb_result skb_frame_get_stats(skb_surface surface, skb_frame_stats* out_stats);

/* Test tap: coverage the last flushed frame computed for raster op `op_index`, as the two
 * planes the fine pass blends in order (directly emitted spans, then accumulated spans), over
 * the rectangle; dst planes are width*height bytes each. */
SKB_API skb_result skb_debug_read_coverage(skb_surface surface, uint32_t op_index, int32_t x, int32_t y,
                                           uint32_t width, uint32_t height, uint8_t* direct, uint8_t* accum);

SKB_API const char* skb_get_last_error_string(void);
SKB_API const char* skb_version_string(void);

#ifdef __cplusplus
}
#endif

#endif /* SKB_H */
