// C entry points over the host-side canvas, for harnesses that are not C++
// (tests/, bench.py): replay an SKSC scene blob (skity_b200/scene.py) through
// CudaCanvas and hand back the encoded display list.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <skity/recorder/display_list.hpp>
#include <skity/recorder/picture_recorder.hpp>

#include "skity_b200/host/cuda_canvas.hpp"
#include "skity_b200/host/scene_player.hpp"
#include "skity_b200/host/skp_player.hpp"

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

// Batch of independent canvases: scene i is replayed onto its own canvas surface of ONE
// display list; surface 0 is a 16x16 dummy.  `scenes` holds the blobs back to back, sizes[i] bytes each.
// canvas_ids[i] receives the surface index of scene i's canvas (blur temporaries take indices in between).
int skbh_encode_scene_batch(const uint8_t* scenes, const size_t* sizes, uint32_t count, uint32_t* canvas_ids,
                            uint8_t** out, size_t* out_n, char* unsupported, size_t cap) {
  skb::DlBuilder builder;
  builder.Reset(16, 16);
  std::string first_unsupported;
  size_t off = 0;
  for (uint32_t i = 0; i < count; i++) {
    if (sizes[i] < sizeof(skb_scene::Header)) return -1;
    skb_scene::Header h;
    std::memcpy(&h, scenes + off, sizeof(h));
    if (h.magic != skb_scene::kMagic) return -1;
    uint32_t sid = builder.AddSurface(h.width, h.height, SKB_SURFACE_CANVAS);
    if (canvas_ids) canvas_ids[i] = sid;
    skity::CudaCanvas canvas(&builder, sid, h.width, h.height);
    int rc = skb_scene::Play(scenes + off, sizes[i], &canvas);
    if (rc != 0) return rc;
    canvas.Flush();
    if (first_unsupported.empty()) first_unsupported = canvas.Unsupported();
    off += sizes[i];
  }
  std::vector<uint8_t> blob = builder.Serialize();
  *out = static_cast<uint8_t*>(std::malloc(blob.size() ? blob.size() : 1));
  if (!*out) return -8;
  std::memcpy(*out, blob.data(), blob.size());
  *out_n = blob.size();
  if (unsupported && cap) {
    std::strncpy(unsupported, first_unsupported.c_str(), cap - 1);
    unsupported[cap - 1] = 0;
  }
  return 0;
}

// The reference's own wire format as input: the scene is first RECORDED with skity::PictureRecorder into a
// skity::DisplayList (include/skity/recorder/display_list.hpp:37-116, src/recorder/recorded_op.hpp) and the
// display list is then replayed onto the CUDA canvas with DisplayList::Draw — the way recorded pictures
// (.skp replays, the golden harness: test/golden/common/golden_test_env.hpp:45-47) reach any backend.
int skbh_encode_recorded_scene(const uint8_t* scene, size_t n, uint8_t** out, size_t* out_n, char* unsupported,
                               size_t cap) {
  if (n < sizeof(skb_scene::Header)) return -1;
  skb_scene::Header h;
  std::memcpy(&h, scene, sizeof(h));
  if (h.magic != skb_scene::kMagic) return -1;
  skity::PictureRecorder recorder;
  recorder.BeginRecording(skity::Rect::MakeWH(static_cast<float>(h.width), static_cast<float>(h.height)));
  int rc = skb_scene::Play(scene, n, recorder.GetRecordingCanvas());
  if (rc != 0) return rc;
  std::unique_ptr<skity::DisplayList> picture = recorder.FinishRecording();
  if (!picture) return -9;
  skb::DlBuilder builder;
  builder.Reset(h.width, h.height);
  skity::CudaCanvas canvas(&builder, 0, h.width, h.height);
  picture->Draw(&canvas);
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

// A serialized picture (.skp, the reference's module/io) as input: read from memory and played back onto the CUDA
// canvas of a width x height surface under the affine matrix m6 (sx kx tx ky sy ty).
int skbh_encode_skp(const uint8_t* skp, size_t n, uint32_t width, uint32_t height, const float* m6, uint8_t** out, size_t* out_n,
                    char* unsupported, size_t cap) {
  if (!skp || !m6 || width == 0 || height == 0) return -1;
  skb::DlBuilder builder;
  builder.Reset(width, height);
  skity::CudaCanvas canvas(&builder, 0, width, height);
  int rc = skb_skp::Play(skp, n, m6, &canvas);
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
