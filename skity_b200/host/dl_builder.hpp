// Host-side builder of the flat display list (include/skb_dl.h).
//
// Deep-copies everything at call time into one arena, the way the reference's
// RecordingCanvas copies Path/Paint into its op stream
// (src/recorder/recorded_op.hpp:238-243): the caller may destroy its Path /
// Paint / Shader as soon as the Canvas call returns.
#ifndef SKITY_B200_HOST_DL_BUILDER_HPP
#define SKITY_B200_HOST_DL_BUILDER_HPP

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "include/skb_dl.h"

namespace skb {

class DlBuilder {
 public:
  DlBuilder() = default;

  void Reset(uint32_t canvas_w, uint32_t canvas_h) {
    surfaces_.clear();
    images_.clear();
    ops_.clear();
    paths_.clear();
    segs_.clear();
    paints_.clear();
    stops_.clear();
    n_clip_states_ = 0;
    AddSurface(canvas_w, canvas_h);
  }

  uint32_t AddSurface(uint32_t w, uint32_t h, uint32_t flags = 0) {
    skb_dl_surface s{};
    s.width = w;
    s.height = h;
    s.flags = flags;
    surfaces_.push_back(s);
    return static_cast<uint32_t>(surfaces_.size() - 1);
  }

  // An application image: RGBA8 pixels (width*height*4 bytes) carried in the display list, stored once however
  // often it is drawn.  An image is identified by its source object AND a hash of the source's bytes at the time
  // of the draw: `keep_alive` (the source pixmap) is held until Reset so that its address cannot be recycled for a
  // different image within the frame, and the hash tells a pixmap that was modified between two draws from itself.
  // FindImageSurface is consulted first so that a repeated draw skips the RGBA conversion.
  static uint64_t HashBytes(const void* p, size_t n) {  // FNV-1a, 8 bytes at a time
    const uint8_t* b = static_cast<const uint8_t*>(p);
    uint64_t h = 1469598103934665603ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
      uint64_t w;
      std::memcpy(&w, b + i, 8);
      h = (h ^ w) * 1099511628211ull;
    }
    for (; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    return h;
  }
  // -> surface id, or 0 (the canvas is never an image) when this (source, content) pair is not in the list yet
  uint32_t FindImageSurface(const void* key, uint64_t content_hash, uint32_t w, uint32_t h) const {
    for (const auto& im : images_)
      if (im.key == key && im.content_hash == content_hash && surfaces_[im.surface].width == w &&
          surfaces_[im.surface].height == h)
        return im.surface;
    return 0;
  }
  uint32_t AddImageSurface(const void* key, uint64_t content_hash, std::shared_ptr<const void> keep_alive, uint32_t w,
                           uint32_t h, const uint8_t* rgba) {
    uint32_t id = AddSurface(w, h, SKB_SURFACE_IMAGE);
    ImageBlob b;
    b.key = key;
    b.content_hash = content_hash;
    b.keep_alive = std::move(keep_alive);
    b.surface = id;
    b.pixels.assign(rgba, rgba + static_cast<size_t>(w) * h * 4);
    images_.push_back(std::move(b));
    return id;
  }

  uint32_t NewClipState() { return ++n_clip_states_; }

  uint32_t AddPath(const std::vector<skb_dl_seg>& segs) {
    skb_dl_path p{};
    p.seg_off = static_cast<uint32_t>(segs_.size());
    p.n_segs = static_cast<uint32_t>(segs.size());
    segs_.insert(segs_.end(), segs.begin(), segs.end());
    paths_.push_back(p);
    return static_cast<uint32_t>(paths_.size() - 1);
  }

  uint32_t AddPaint(const skb_dl_paint& p) {
    paints_.push_back(p);
    return static_cast<uint32_t>(paints_.size() - 1);
  }

  // returns the float offset of the block in the stop pool
  uint32_t AddStops(const float* colors4, const float* stops, uint32_t n) {
    uint32_t off = static_cast<uint32_t>(stops_.size());
    stops_.insert(stops_.end(), colors4, colors4 + 4 * static_cast<size_t>(n));
    if (stops) {
      stops_.insert(stops_.end(), stops, stops + n);
    } else {
      stops_.insert(stops_.end(), n, 0.f);
    }
    return off;
  }

  // appends to the float pool (directly behind the block AddStops just returned)
  void AddFloats(const float* v, uint32_t n) { stops_.insert(stops_.end(), v, v + n); }

  // raw 32-bit words in the float pool (colour-filter blocks); returns their word offset
  uint32_t AddWords(const uint32_t* v, uint32_t n) {
    uint32_t off = static_cast<uint32_t>(stops_.size());
    stops_.resize(stops_.size() + n);
    std::memcpy(stops_.data() + off, v, 4 * static_cast<size_t>(n));
    return off;
  }

  void AddOp(const skb_dl_op& op) { ops_.push_back(op); }

  size_t OpCount() const { return ops_.size(); }
  uint32_t SurfaceCount() const { return static_cast<uint32_t>(surfaces_.size()); }

  std::vector<uint8_t> Serialize() const {
    auto align16 = [](size_t v) { return (v + 15) & ~static_cast<size_t>(15); };
    skb_dl_header h{};
    h.magic = SKB_DL_MAGIC;
    h.version = SKB_DL_VERSION;
    h.n_surfaces = static_cast<uint32_t>(surfaces_.size());
    h.n_ops = static_cast<uint32_t>(ops_.size());
    h.n_paths = static_cast<uint32_t>(paths_.size());
    h.n_segs = static_cast<uint32_t>(segs_.size());
    h.n_paints = static_cast<uint32_t>(paints_.size());
    h.n_stop_floats = static_cast<uint32_t>(stops_.size());
    h.n_clip_states = n_clip_states_;
    size_t off = align16(sizeof(h));
    h.off_surfaces = static_cast<uint32_t>(off);
    off = align16(off + surfaces_.size() * sizeof(skb_dl_surface));
    h.off_ops = static_cast<uint32_t>(off);
    off = align16(off + ops_.size() * sizeof(skb_dl_op));
    h.off_paths = static_cast<uint32_t>(off);
    off = align16(off + paths_.size() * sizeof(skb_dl_path));
    h.off_segs = static_cast<uint32_t>(off);
    off = align16(off + segs_.size() * sizeof(skb_dl_seg));
    h.off_paints = static_cast<uint32_t>(off);
    off = align16(off + paints_.size() * sizeof(skb_dl_paint));
    h.off_stops = static_cast<uint32_t>(off);
    off = align16(off + stops_.size() * sizeof(float));
    std::vector<skb_dl_surface> surfaces = surfaces_;
    for (const auto& im : images_) {  // image pixels go last
      surfaces[im.surface].reserved = static_cast<uint32_t>(off);
      off = align16(off + im.pixels.size());
    }
    h.total_bytes = static_cast<uint32_t>(off);
    std::vector<uint8_t> out(off, 0);
    std::memcpy(out.data(), &h, sizeof(h));
    auto put = [&](uint32_t o, const void* p, size_t n) {
      if (n) std::memcpy(out.data() + o, p, n);
    };
    put(h.off_surfaces, surfaces.data(), surfaces.size() * sizeof(skb_dl_surface));
    for (const auto& im : images_) put(surfaces[im.surface].reserved, im.pixels.data(), im.pixels.size());
    put(h.off_ops, ops_.data(), ops_.size() * sizeof(skb_dl_op));
    put(h.off_paths, paths_.data(), paths_.size() * sizeof(skb_dl_path));
    put(h.off_segs, segs_.data(), segs_.size() * sizeof(skb_dl_seg));
    put(h.off_paints, paints_.data(), paints_.size() * sizeof(skb_dl_paint));
    put(h.off_stops, stops_.data(), stops_.size() * sizeof(float));
    return out;
  }

 private:
  struct ImageBlob {
    const void* key;
    uint64_t content_hash;
    std::shared_ptr<const void> keep_alive;
    uint32_t surface;
    std::vector<uint8_t> pixels;
  };
  std::vector<skb_dl_surface> surfaces_;
  std::vector<ImageBlob> images_;
  std::vector<skb_dl_op> ops_;
  std::vector<skb_dl_path> paths_;
  std::vector<skb_dl_seg> segs_;
  std::vector<skb_dl_paint> paints_;
  std::vector<float> stops_;
  uint32_t n_clip_states_ = 0;
};

}  // namespace skb

#endif  // SKITY_B200_HOST_DL_BUILDER_HPP
