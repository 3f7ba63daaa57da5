// Public header of the CUDA backend plug-in, the counterpart of the reference's
// include/skity/gpu/gpu_context_gl.hpp (GLContextCreate, :115) for this backend.
//
//   auto ctx  = skity::CudaContextCreate(&desc);                  // std::unique_ptr<GPUContext>
//   skity::GPUSurfaceDescriptorCuda sd; sd.backend = skity::kGPUBackendTypeCUDA; sd.width = ...;
//   auto surf = ctx->CreateSurface(&sd);                          // std::unique_ptr<GPUSurface>
//   skity::Canvas* canvas = surf->LockCanvas(true);               // same Canvas API as every backend
//   canvas->DrawPath(path, paint); ...
//   surf->Flush();                                                // encodes the frame, launches the CUDA stages
//   auto pixmap = surf->ReadPixels(skity::Rect::MakeWH(w, h));    // premultiplied RGBA8
#ifndef SKITY_B200_HOST_GPU_CONTEXT_CUDA_HPP
#define SKITY_B200_HOST_GPU_CONTEXT_CUDA_HPP

#include <memory>
#include <skity/gpu/gpu_context.hpp>
#include <skity/gpu/gpu_surface.hpp>

namespace skity {

// GPUBackendType (include/skity/gpu/gpu_backend_type.hpp:13-44) has no CUDA enumerator; the value
// lives here instead of editing the reference's header.
constexpr GPUBackendType kGPUBackendTypeCUDA = static_cast<GPUBackendType>(100);

struct CudaContextDesc {
  int device_ordinal = 0;
};

// Descriptor subclass, checked through desc->backend like GPUSurfaceDescriptorGL
// (src/gpu/gl/gpu_context_impl_gl.cc:87-93).
struct GPUSurfaceDescriptorCuda : public GPUSurfaceDescriptor {
  // Tile band [band_y0, band_y1) this device renders (rows, multiples of 16); 0,0 = whole surface.
  uint32_t band_y0 = 0;
  uint32_t band_y1 = 0;
};

// Returns null when no sm_100-class CUDA device is usable (there is no CPU fallback).
SKITY_API std::unique_ptr<GPUContext> CudaContextCreate(const CudaContextDesc* desc);

}  // namespace skity

#endif  // SKITY_B200_HOST_GPU_CONTEXT_CUDA_HPP
