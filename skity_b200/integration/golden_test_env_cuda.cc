// Harness-side integration source — compiled inside the REFERENCE's golden-test harness (test/golden), not by this
// repo's build: the CUDA counterpart of test/golden/common/{gl,mtl,vk}/golden_test_env_*.cc.
//
// A maintainer adds, in the reference tree:
//   test/golden/common/golden_test_env.hpp:21-25   enum class Backend { kGL, kVulkan, kMetal, kCUDA };
//   test/golden/common/golden_test_env.cc:15-41    GoldenTestEnv* CreateGoldenTestEnvCUDA();  and, in CreateInstance,
//                                                  case Backend::kCUDA: g_golden_test_env = CreateGoldenTestEnvCUDA(); break;
//   test/golden/CMakeLists.txt                     this file + target_link_libraries(... skb_skity skb)
// after which every golden case (DisplayListToTexture -> RenderToTexture, golden_test_env.hpp:37-47) renders through
// CudaContextCreate -> CreateSurface -> LockCanvas -> DisplayList::Draw -> Flush -> ReadPixels, and
// CompareGoldenTexture (common/golden_test_check.hpp) compares the pixels with the checked-in images under the
// software backend's tolerance.
//
// Syntax-checked here against the reference's own harness headers with a two-line stand-in for <gtest/gtest.h>
// (tests/test_host_and_abi.py::test_golden_harness_env_compiles_against_the_reference_headers); gtest itself is not
// in this image, so the harness cannot be built or run here.
#include <skity/skity.hpp>

#include "common/golden_test_env.hpp"
#include "common/golden_texture.hpp"
#include "skity_b200/host/gpu_context_cuda.hpp"

namespace skity {
namespace testing {

namespace {

// The frame is already on the host when the texture object is made: ReadPixels hands it out.
class GoldenTextureCUDA : public GoldenTexture {
 public:
  explicit GoldenTextureCUDA(std::shared_ptr<Pixmap> pixels) : GoldenTexture(Image::MakeImage(pixels)), pixels_(std::move(pixels)) {}
  std::shared_ptr<Pixmap> ReadPixels() override { return pixels_; }

 private:
  std::shared_ptr<Pixmap> pixels_;
};

class GoldenTestEnvCUDA : public GoldenTestEnv {
 public:
  Backend GetBackend() const override { return static_cast<Backend>(3); }   // Backend::kCUDA once the enumerator exists

  std::shared_ptr<GoldenTexture> RenderToTexture(uint32_t width, uint32_t height,
                                                 const std::function<void(Canvas*)>& render) override {
    GPUSurfaceDescriptorCuda desc;
    desc.backend = kGPUBackendTypeCUDA;
    desc.width = width;
    desc.height = height;
    desc.content_scale = 1.f;
    desc.sample_count = 1;   // coverage is analytic: no MSAA resolve
    std::unique_ptr<GPUSurface> surface = GetGPUContext()->CreateSurface(&desc);
    if (!surface) return nullptr;
    Canvas* canvas = surface->LockCanvas(true);
    render(canvas);
    canvas->Flush();
    surface->Flush();
    std::shared_ptr<Pixmap> pixels = surface->ReadPixels(Rect::MakeWH(static_cast<float>(width), static_cast<float>(height)));
    if (!pixels) return nullptr;
    return std::make_shared<GoldenTextureCUDA>(std::move(pixels));
  }

 protected:
  std::unique_ptr<GPUContext> CreateGPUContext() override {
    CudaContextDesc desc;
    desc.device_ordinal = 0;
    return CudaContextCreate(&desc);   // null without an sm_100-class GPU: SetUp leaves gpu_context_ empty, cases fail loudly
  }
};

}  // namespace

GoldenTestEnv* CreateGoldenTestEnvCUDA() { return new GoldenTestEnvCUDA(); }

}  // namespace testing
}  // namespace skity
