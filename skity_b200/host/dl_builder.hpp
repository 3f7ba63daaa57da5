// Host-side builder of the flat display list (include/skb_dl.h).
//
// Deep-copies everything at call time into one arena, the way the reference's
// RecordingCanvas copies Path/Paint into its op stream
// (src/recorder/recorded_op.hpp:238-243): the caller may destroy its Path /
// Paint / Shader as soon as the Canvas call returns.
#ifndef SKITY_B200_HOST_DL_BUILDER_HPP
#define SKITY_B200_HOST_DL_BUILDER_HPP

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "include/skb_dl.h"

namespace skb {

class DlBuilder {
 public:
  DlBuilder() = default;
  DlBuilder(const DlBuilder&) = delete;
  DlBuilder& operator=(const DlBuilder&) = delete;
  ~DlBuilder() { StopWorkers(); }

  void Reset(uint32_t canvas_w, uint32_t canvas_h) {
    WaitForJobs();
    deferred_.clear();
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

  // A path whose segments are worked out later, by `job`, on one of the builder's worker threads: the outline of a
  // stroke (Stroke::StrokePath is by far the most expensive thing a Canvas call does on the host, and one draw's
  // outline does not depend on another's).  The op that owns the path is added right away, so draw order is kept;
  // Finish() waits for the jobs and drops the ops whose path came out empty — what the immediate route does by not
  // adding them in the first place.  `job` must own copies of everything it reads.
  using SegJob = std::function<void(std::vector<skb_dl_seg>*)>;
  uint32_t AddDeferredPath(SegJob job) {
    deferred_.emplace_back();
    std::vector<skb_dl_seg>* out = &deferred_.back();   // a deque: the address stays valid while more are added
    skb_dl_path p{};
    p.seg_off = static_cast<uint32_t>(segs_.size());
    p.n_segs = 0;
    p.reserved[0] = static_cast<uint32_t>(deferred_.size());   // 1 + index into deferred_
    paths_.push_back(p);
    static const unsigned want = [] {   // SKB_HOST_THREADS=1: outlines worked out by the calling thread
      const char* e = std::getenv("SKB_HOST_THREADS");
      const unsigned n = e ? static_cast<unsigned>(std::atoi(e)) : std::thread::hardware_concurrency();
      return std::max(1u, std::min(16u, n));
    }();
    if (want <= 1) {
      job(out);
    } else {
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (workers_.empty())
          for (unsigned t = 0; t < want; t++) workers_.emplace_back([this] { WorkerLoop(); });
        jobs_.emplace_back([job = std::move(job), out] { job(out); });
        pending_++;
      }
      cv_work_.notify_one();
    }
    return static_cast<uint32_t>(paths_.size() - 1);
  }

  // Waits for the deferred paths and folds them into the flat tables (every fill / clip op owns the next path, the
  // paths tile the segment table).  Must be called before MakeLayout / Serialize when AddDeferredPath was used; cheap
  // otherwise.
  void Finish() {
    WaitForJobs();
    if (deferred_.empty()) return;
    std::vector<skb_dl_op> ops;
    std::vector<skb_dl_path> paths;
    std::vector<skb_dl_seg> segs;
    ops.reserve(ops_.size());
    paths.reserve(paths_.size());
    size_t total = segs_.size();
    for (const auto& d : deferred_) total += d.size();
    segs.reserve(total);
    for (skb_dl_op o : ops_) {
      if (o.kind == SKB_OP_FILL || o.kind == SKB_OP_CLIP) {
        const skb_dl_path& p = paths_[o.path];
        const skb_dl_seg* src = segs_.data() + p.seg_off;
        size_t n = p.n_segs;
        if (p.reserved[0]) {
          const std::vector<skb_dl_seg>& d = deferred_[p.reserved[0] - 1];
          if (d.empty()) continue;   // nothing to fill: the op is not part of the list
          src = d.data();
          n = d.size();
        }
        skb_dl_path np{};
        np.seg_off = static_cast<uint32_t>(segs.size());
        np.n_segs = static_cast<uint32_t>(n);
        segs.insert(segs.end(), src, src + n);
        o.path = static_cast<uint32_t>(paths.size());
        paths.push_back(np);
      }
      ops.push_back(o);
    }
    ops_.swap(ops);
    paths_.swap(paths);
    segs_.swap(segs);
    deferred_.clear();
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

  // The flat display list in two steps, so that the caller can hand in its own (page-locked) buffer:
  // Layout() fixes the section offsets and the total size, SerializeInto() writes `total_bytes` bytes.
  struct Layout {
    skb_dl_header h;
    std::vector<uint32_t> image_off;   // byte offset of every image's pixels
  };
  Layout MakeLayout() const {
    auto align16 = [](size_t v) { return (v + 15) & ~static_cast<size_t>(15); };
    Layout L;
    skb_dl_header& h = L.h;
    h = skb_dl_header{};
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
    for (const auto& im : images_) {  // image pixels go last
      L.image_off.push_back(static_cast<uint32_t>(off));
      off = align16(off + im.pixels.size());
    }
    h.total_bytes = static_cast<uint32_t>(off);
    return L;
  }

  // `out`: L.h.total_bytes bytes.  Every byte is written (padding as zero), so a reused buffer needs no clearing.  The
  // large sections are copied by a few threads: at 1M draws the list is ~0.5 GB and one thread's memcpy would cost more
  // than the frame takes on the device.
  void SerializeInto(const Layout& L, uint8_t* out) const {
    const skb_dl_header& h = L.h;
    std::vector<skb_dl_surface> surfaces = surfaces_;
    for (size_t i = 0; i < images_.size(); i++) surfaces[images_[i].surface].reserved = L.image_off[i];
    struct Part { size_t off; const void* p; size_t n; };
    std::vector<Part> parts;
    parts.push_back({0, &h, sizeof(h)});
    parts.push_back({h.off_surfaces, surfaces.data(), surfaces.size() * sizeof(skb_dl_surface)});
    parts.push_back({h.off_ops, ops_.data(), ops_.size() * sizeof(skb_dl_op)});
    parts.push_back({h.off_paths, paths_.data(), paths_.size() * sizeof(skb_dl_path)});
    parts.push_back({h.off_segs, segs_.data(), segs_.size() * sizeof(skb_dl_seg)});
    parts.push_back({h.off_paints, paints_.data(), paints_.size() * sizeof(skb_dl_paint)});
    parts.push_back({h.off_stops, stops_.data(), stops_.size() * sizeof(float)});
    for (size_t i = 0; i < images_.size(); i++) parts.push_back({L.image_off[i], images_[i].pixels.data(), images_[i].pixels.size()});
    // the gaps between the parts (alignment padding, at most 15 bytes each) are zeroed
    for (size_t i = 0; i < parts.size(); i++) {
      const size_t end = parts[i].off + parts[i].n;
      const size_t next = i + 1 < parts.size() ? parts[i + 1].off : h.total_bytes;
      if (next > end) std::memset(out + end, 0, next - end);
    }
    // work items of at most 8 MB, dealt to the threads round-robin
    struct Item { uint8_t* d; const uint8_t* s; size_t n; };
    std::vector<Item> items;
    const size_t kChunk = static_cast<size_t>(8) << 20;
    for (const Part& p : parts)
      for (size_t o = 0; o < p.n; o += kChunk)
        items.push_back({out + p.off + o, static_cast<const uint8_t*>(p.p) + o, std::min(kChunk, p.n - o)});
    unsigned nt = h.total_bytes >= (static_cast<size_t>(32) << 20) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)) : 1u;
    auto work = [&](unsigned t) {
      for (size_t i = t; i < items.size(); i += nt) std::memcpy(items[i].d, items[i].s, items[i].n);
    };
    if (nt <= 1) {
      work(0);
      return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }

  std::vector<uint8_t> Serialize() {
    Finish();
    const Layout L = MakeLayout();
    std::vector<uint8_t> out(L.h.total_bytes);
    SerializeInto(L, out.data());
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

  // deferred paths and the threads that work them out
  void WorkerLoop() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
        if (jobs_.empty()) return;   // stop_
        job = std::move(jobs_.front());
        jobs_.pop_front();
      }
      job();
      {
        std::lock_guard<std::mutex> lk(mu_);
        pending_--;
      }
      cv_done_.notify_all();
    }
  }
  void WaitForJobs() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return pending_ == 0; });
  }
  void StopWorkers() {
    WaitForJobs();
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : workers_) t.join();
    workers_.clear();
  }
  std::deque<std::vector<skb_dl_seg>> deferred_;
  std::deque<std::function<void()>> jobs_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  size_t pending_ = 0;
  bool stop_ = false;
};

}  // namespace skb

#endif  // SKITY_B200_HOST_DL_BUILDER_HPP
