// libcompactb200: ABI bookkeeping (version, errors, workspace sizing) and the host-buffer
// entry points that wrap H2D copy + kernels + D2H copy (the "e2e" path of bench.py).
#include <stdarg.h>
#include <string.h>

#include "cf_common.cuh"

namespace cf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

size_t sign_codec_workspace_bytes(int64_t N, int64_t C, int batch);
size_t minmax_codec_workspace_bytes(int64_t N, int64_t C);
size_t lowrank_workspace_bytes(int64_t N, int64_t C, int rank);

// wire payload geometry: [codes | scale_u (N) | scale_v (C)] in bytes (main.py:149-152)
struct PayloadLayout {
  size_t code_bytes, u_off, v_off, total;
};
static PayloadLayout payload_layout(int codec, int64_t N, int64_t C) {
  PayloadLayout l{};
  const size_t e = static_cast<size_t>(N) * C;
  if (codec == CF_CODEC_BINARY) l.code_bytes = e / 8;
  else if (codec == CF_CODEC_INT2) l.code_bytes = e / 4;
  l.u_off = l.code_bytes;
  l.v_off = l.u_off + static_cast<size_t>(N) * 2;
  l.total = l.v_off + static_cast<size_t>(C) * 2;
  return l;
}

}  // namespace cf

extern "C" {

int cf_abi_version(void) { return CF_ABI_VERSION; }
const char* cf_last_error(void) { return cf::g_err; }

int cf_sm_count(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cf::set_error("cudaGetDevice failed (no CUDA device?)");
    return CF_ERR_CUDA;
  }
  return cf::sm_count();
}

size_t cf_workspace_bytes(int codec, int64_t N, int64_t C, int rank, int batch) {
  if (N <= 0 || C <= 0) return 0;
  switch (codec) {
    case CF_CODEC_BINARY:
    case CF_CODEC_INT2:
      return cf::sign_codec_workspace_bytes(N, C, batch);
    case CF_CODEC_INT4:
    case CF_CODEC_INT8:
      return cf::minmax_codec_workspace_bytes(N, C) * static_cast<size_t>(batch > 0 ? batch : 1);
    case CF_CODEC_LOWRANK:
      return cf::lowrank_workspace_bytes(N, C, rank) * static_cast<size_t>(batch > 0 ? batch : 1);
    default:
      return 256;
  }
}

size_t cf_host_scratch_bytes(int codec, int64_t N, int64_t C) {
  if (codec != CF_CODEC_BINARY && codec != CF_CODEC_INT2) return 0;
  const size_t e2 = cf::round_up(static_cast<size_t>(N) * C * 2, 256);
  const cf::PayloadLayout l = cf::payload_layout(codec, N, C);
  // x | base | new_base/recon | payload | workspace
  return 3 * e2 + cf::round_up(l.total, 256) + cf_workspace_bytes(codec, N, C, 0, 1);
}

int cf_host_compress(int codec, const void* x_host, const void* base_host, void* new_base_host,
                     void* payload_host, int64_t N, int64_t C, void* dev_scratch, size_t dev_scratch_bytes,
                     cf_stream_t stream) {
  CF_CHECK_ARG(codec == CF_CODEC_BINARY || codec == CF_CODEC_INT2, "host entry points support BINARY and INT2");
  CF_CHECK_ARG(x_host && payload_host && dev_scratch, "null pointer");
  CF_CHECK_ARG(N >= 1 && C >= 8 && C % 8 == 0, "bad shape");
  CF_CHECK_ARG(dev_scratch_bytes >= cf_host_scratch_bytes(codec, N, C), "device scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = static_cast<size_t>(N) * C * 2;
  const size_t e2 = cf::round_up(bytes, 256);
  const cf::PayloadLayout l = cf::payload_layout(codec, N, C);
  char* d = static_cast<char*>(dev_scratch);
  char* dx = d; char* dbase = d + e2; char* dnew = d + 2 * e2; char* dpay = d + 3 * e2;
  char* dws = dpay + cf::round_up(l.total, 256);
  const size_t ws_bytes = cf_workspace_bytes(codec, N, C, 0, 1);
  CF_CHECK_CUDA(cudaMemcpyAsync(dx, x_host, bytes, cudaMemcpyHostToDevice, st));
  if (base_host) CF_CHECK_CUDA(cudaMemcpyAsync(dbase, base_host, bytes, cudaMemcpyHostToDevice, st));
  int rc;
  if (codec == CF_CODEC_BINARY)
    rc = cf_binary_compress(dx, base_host ? dbase : nullptr, new_base_host ? dnew : nullptr, dpay, dpay + l.u_off,
                            dpay + l.v_off, N, C, dws, ws_bytes, stream);
  else
    rc = cf_int2_compress(dx, base_host ? dbase : nullptr, new_base_host ? dnew : nullptr, dpay, dpay + l.u_off,
                          dpay + l.v_off, N, C, dws, ws_bytes, stream);
  if (rc) return rc;
  CF_CHECK_CUDA(cudaMemcpyAsync(payload_host, dpay, l.total, cudaMemcpyDeviceToHost, st));
  if (new_base_host) CF_CHECK_CUDA(cudaMemcpyAsync(new_base_host, dnew, bytes, cudaMemcpyDeviceToHost, st));
  CF_CHECK_CUDA(cudaStreamSynchronize(st));
  return CF_OK;
}

int cf_host_decompress(int codec, const void* payload_host, const void* base_host, void* recon_host, int64_t N,
                       int64_t C, void* dev_scratch, size_t dev_scratch_bytes, cf_stream_t stream) {
  CF_CHECK_ARG(codec == CF_CODEC_BINARY || codec == CF_CODEC_INT2, "host entry points support BINARY and INT2");
  CF_CHECK_ARG(payload_host && recon_host && dev_scratch, "null pointer");
  CF_CHECK_ARG(N >= 1 && C >= 8 && C % 8 == 0, "bad shape");
  CF_CHECK_ARG(dev_scratch_bytes >= cf_host_scratch_bytes(codec, N, C), "device scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = static_cast<size_t>(N) * C * 2;
  const size_t e2 = cf::round_up(bytes, 256);
  const cf::PayloadLayout l = cf::payload_layout(codec, N, C);
  char* d = static_cast<char*>(dev_scratch);
  char* dbase = d + e2; char* drec = d + 2 * e2; char* dpay = d + 3 * e2;
  CF_CHECK_CUDA(cudaMemcpyAsync(dpay, payload_host, l.total, cudaMemcpyHostToDevice, st));
  if (base_host) CF_CHECK_CUDA(cudaMemcpyAsync(dbase, base_host, bytes, cudaMemcpyHostToDevice, st));
  int rc;
  if (codec == CF_CODEC_BINARY)
    rc = cf_binary_decompress(dpay, dpay + l.u_off, dpay + l.v_off, 1, base_host ? dbase : nullptr, drec, N, C, stream);
  else
    rc = cf_int2_decompress(dpay, dpay + l.u_off, dpay + l.v_off, base_host ? dbase : nullptr, drec, N, C, stream);
  if (rc) return rc;
  CF_CHECK_CUDA(cudaMemcpyAsync(recon_host, drec, bytes, cudaMemcpyDeviceToHost, st));
  CF_CHECK_CUDA(cudaStreamSynchronize(st));
  return CF_OK;
}

}  // extern "C"
