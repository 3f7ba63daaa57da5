// C entry points over the host-side canvas, for harnesses that are not C++
// (tests/, bench.py): replay an SKSC scene blob (skity_b200/scene.py) through
// CudaCanvas and hand back the encoded display list.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "skity_b200/host/cuda_canvas.hpp"
#include "skity_b200/host/scene_player.hpp"

extern "C" {

// Returns 0 on success; *out is malloc'd (free with skbh_free) and holds *out_n bytes.
// `unsupported` (optional, cap bytes) receives the first feature the backend had to drop.
int skbh_encode_scene(const uint8_t* scene, size_t n, uint8_t** out, size_t* out_n, char* unsupported,
                      size_t cap) {
  if (n < sizeof(skb_scene::Header)) return -1;
  skb_scene::Header h;
  std::memcpy(&h, scene, sizeof(h));
  if (h.magic != skb_scene::kMagic) return -1;
  skb::DlBuilder builder;
  builder.Reset(h.width, h.height);
  skity::CudaCanvas canvas(&builder, 0, h.width, h.height);
  int rc = skb_scene::Play(scene, n, &canvas);
  if (rc != 0) return rc;
  canvas.Flush();
  std::vector<uint8_t> blob = builder.Serialize();
  *out = static_cast<uint8_t*>(std::malloc(blob.size() ? blob.size() : 1));
  if (!*out) return -8;
  std::memcpy(*out, blob.data(), blob.size());
  *out_n = blob.size();
  if (unsupported && cap) {
    std::strncpy(unsupported, canvas.Unsupported().c_str(), cap - 1);
    unsupported[cap - 1] = 0;
  }
  return 0;
}

void skbh_free(void* p) { std::free(p); }

}  // extern "C"
