// jpeg.cu -- JPEG decoding in front of the batch-assembly kernel ("nvJPEG ... + fused resize / normalise
// kernel"): dataset.py:20-24 `Image.open(file).convert('RGB')` for JPEG files, executed by nvJPEG (a CUDA-toolkit LIBRARY, not
// our code -- calling it counts like calling cuBLAS) straight into the device [B,H,W,C] uint8 batch that k_image_batch reads.
// Entropy (Huffman) decoding is bit-serial and runs on the host inside nvJPEG's hybrid backend; IDCT, chroma upsampling and
// colour conversion run on the GPU, in stream order.
//
// NOT bit-exact with Pillow: libjpeg-turbo (Pillow's decoder) and nvJPEG round the IDCT / YCbCr->RGB conversion differently
// and libjpeg applies "fancy" (triangle-filter) chroma upsampling to 4:2:0 files.  The default loader therefore keeps decoding
// with Pillow on the host (bit-exact with the reference); this path is opt-in (SIVAE_GPU_JPEG=1, gpu_dataset.py).
//
// nvJPEG is resolved at run time (dlopen) so that the library has no link-time dependency on it.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>

#include <cuda_runtime.h>
#include <nvjpeg.h>      // types and enums only: the functions are resolved with dlsym

#include "kernels.h"

namespace sivae {

namespace {
struct NvjpegApi {
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  std::string error;
  bool ok = false;
};
std::mutex g_jpeg_mu;

NvjpegApi* jpeg_api() {            // call with g_jpeg_mu held
  static NvjpegApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {getenv("SIVAE_NVJPEG_LIB"), "libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    if (!nm || !nm[0]) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) { api.error = "libnvjpeg.so.12 not found (set SIVAE_NVJPEG_LIB)"; return &api; }
  api.CreateSimple = (decltype(api.CreateSimple))dlsym(h, "nvjpegCreateSimple");
  api.Destroy = (decltype(api.Destroy))dlsym(h, "nvjpegDestroy");
  api.JpegStateCreate = (decltype(api.JpegStateCreate))dlsym(h, "nvjpegJpegStateCreate");
  api.JpegStateDestroy = (decltype(api.JpegStateDestroy))dlsym(h, "nvjpegJpegStateDestroy");
  api.GetImageInfo = (decltype(api.GetImageInfo))dlsym(h, "nvjpegGetImageInfo");
  api.Decode = (decltype(api.Decode))dlsym(h, "nvjpegDecode");
  if (!api.CreateSimple || !api.Destroy || !api.JpegStateCreate || !api.JpegStateDestroy || !api.GetImageInfo || !api.Decode) {
    api.error = "libnvjpeg lacks the nvjpegCreateSimple / nvjpegDecode entry points";
    return &api;
  }
  nvjpegStatus_t r = api.CreateSimple(&api.handle);
  if (r == NVJPEG_STATUS_SUCCESS) r = api.JpegStateCreate(api.handle, &api.state);
  if (r != NVJPEG_STATUS_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof(buf), "nvjpegCreateSimple / nvjpegJpegStateCreate failed (nvjpegStatus_t %d)", (int)r);
    api.error = buf;
    return &api;
  }
  api.ok = true;
  return &api;
}
}  // namespace

// 0 = ok; -9 = nvJPEG unavailable; -8 = an image is not a baseline JPEG nvJPEG decodes / its size or component count differs
// from (height, width, 3 components); > 0 = cudaError_t.  *msg: description of the failure.
// data[i] / lengths[i]: HOST pointers to the compressed files; out_hwc: DEVICE [batch][height][width][3] (RGB, interleaved).
int jpeg_decode_batch(const unsigned char* const* data, const long long* lengths, int batch, int height, int width,
                      unsigned char* out_hwc, cudaStream_t st, std::string* msg) {
  std::lock_guard<std::mutex> lk(g_jpeg_mu);
  NvjpegApi* a = jpeg_api();
  if (!a->ok) { *msg = a->error; return -9; }
  const int C = 3;
  for (int i = 0; i < batch; ++i) {
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0}, hs[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0};
    nvjpegChromaSubsampling_t css;
    nvjpegStatus_t r = a->GetImageInfo(a->handle, data[i], (size_t)lengths[i], &ncomp, &css, ws, hs);
    if (r != NVJPEG_STATUS_SUCCESS || ncomp != 3 || ws[0] != width || hs[0] != height) {
      char buf[160];
      snprintf(buf, sizeof(buf), "image %d: nvjpegGetImageInfo status %d, %d components, %d x %d (expected 3 components, %d x %d)", i,
               (int)r, ncomp, ws[0], hs[0], width, height);
      *msg = buf;
      return -8;
    }
    nvjpegImage_t img;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { img.channel[c] = nullptr; img.pitch[c] = 0; }
    img.channel[0] = out_hwc + (size_t)i * height * width * C;
    img.pitch[0] = (size_t)width * C;
    // RGBI: interleaved RGB into channel 0 = the [H,W,3] layout of a decoded PIL image
    r = a->Decode(a->handle, a->state, data[i], (size_t)lengths[i], NVJPEG_OUTPUT_RGBI, &img, st);
    if (r != NVJPEG_STATUS_SUCCESS) {
      char buf[96];
      snprintf(buf, sizeof(buf), "image %d: nvjpegDecode failed (nvjpegStatus_t %d)", i, (int)r);
      *msg = buf;
      return -8;
    }
  }
  // the compressed bytes belong to the caller: do not return while nvJPEG may still read them
  cudaError_t ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) { *msg = cudaGetErrorString(ce); return (int)ce; }
  return 0;
}

// header only (no CUDA work): height, width, components of a JPEG file; -8 / -9 as above
int jpeg_info(const unsigned char* data, long long length, int* height, int* width, int* components, std::string* msg) {
  std::lock_guard<std::mutex> lk(g_jpeg_mu);
  NvjpegApi* a = jpeg_api();
  if (!a->ok) { *msg = a->error; return -9; }
  int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0}, hs[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0};
  nvjpegChromaSubsampling_t css;
  nvjpegStatus_t r = a->GetImageInfo(a->handle, data, (size_t)length, &ncomp, &css, ws, hs);
  if (r != NVJPEG_STATUS_SUCCESS) { *msg = "nvjpegGetImageInfo failed (not a JPEG stream nvJPEG parses)"; return -8; }
  *height = hs[0]; *width = ws[0]; *components = ncomp;
  return 0;
}

}  // namespace sivae
