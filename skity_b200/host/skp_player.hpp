// .skp ingest: a serialized picture (the reference's module/io, Skia's SKP format up to version 109) read from memory
// and played back onto any skity::Canvas — the CUDA canvas in the plug-in, the software canvas in the oracle build.
// This is how the reference's own SKP golden case and benchmark feed a backend (test/golden/cases/skp/skp.cc:51-68,
// test/bench/case/draw_skp.cc:12-29): ReadStream -> Picture::MakeFromStream -> Picture::PlayBack(canvas).
#ifndef SKITY_B200_HOST_SKP_PLAYER_HPP
#define SKITY_B200_HOST_SKP_PLAYER_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
// picture.hpp default-initialises a std::unique_ptr<MemoryWriter32> member in the class body, which g++ 13 only accepts
// where that type is complete: the module's own header first (module/io/src/io/memory_writer.hpp)
#include "src/io/memory_writer.hpp"

#include <skity/io/picture.hpp>
#include <skity/io/stream.hpp>
#include <skity/render/canvas.hpp>

namespace skb_skp {

// skity::ReadStream (module/io/include/skity/io/stream.hpp:47-113) over a caller-owned byte range.
class MemoryReadStream final : public skity::ReadStream {
 public:
  MemoryReadStream(const uint8_t* p, size_t n) : p_(p), n_(n) {}
  size_t Read(void* buffer, size_t size) override {
    const size_t k = std::min(size, n_ - pos_);
    if (buffer && k) std::memcpy(buffer, p_ + pos_, k);
    pos_ += k;
    return k;
  }
  size_t Peek(void* buffer, size_t size) override {
    const size_t k = std::min(size, n_ - pos_);
    if (buffer && k) std::memcpy(buffer, p_ + pos_, k);
    return k;
  }
  bool IsAtEnd() const override { return pos_ >= n_; }
  bool Rewind() override {
    pos_ = 0;
    return true;
  }

 private:
  const uint8_t* p_;
  size_t n_;
  size_t pos_ = 0;
};

// Plays the picture under the affine matrix m6 = sx kx tx ky sy ty.  0 on success, -20: not a readable picture.
inline int Play(const uint8_t* skp, size_t n, const float* m6, skity::Canvas* canvas) {
  MemoryReadStream stream(skp, n);
  auto picture = skity::Picture::MakeFromStream(stream);
  if (!picture) return -20;
  canvas->Save();
  canvas->Concat(skity::Matrix(m6[0], m6[1], m6[2], m6[3], m6[4], m6[5], 0.f, 0.f, 1.f));   // row-major 2x3, as scene_player's Affine()
  picture->PlayBack(canvas);
  canvas->Restore();
  return 0;
}

}  // namespace skb_skp

#endif  // SKITY_B200_HOST_SKP_PLAYER_HPP
