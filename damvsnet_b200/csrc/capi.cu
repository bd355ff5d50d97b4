// C-ABI glue: error state, launch counter, device check, conv dispatch.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace damvs {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// conv3d_direct.cu
int conv3d_direct_launch(const damvs_conv3d_desc*, const void*, const void*, const float*, const float*, const void*,
                         void*, cudaStream_t);
size_t conv3d_direct_packed_bytes(const damvs_conv3d_desc*);
int conv3d_direct_pack(const damvs_conv3d_desc*, const float*, void*, cudaStream_t);
// conv3d_tc.cu
int conv3d_tc_launch(const damvs_conv3d_desc*, const void*, const void*, const float*, const float*, const void*,
                     void*, cudaStream_t);
size_t conv3d_tc_packed_bytes(const damvs_conv3d_desc*);
int conv3d_tc_pack(const damvs_conv3d_desc*, const float*, void*, cudaStream_t);

bool conv3d_tcp_head_supported(const damvs_conv3d_desc*);
int conv3d_tcp_head_launch(const damvs_conv3d_desc*, const void*, const void*, const float*, float*, float*, float*, float*, cudaStream_t);

static int check_conv_desc(const damvs_conv3d_desc* d) {
  DAMVS_REQUIRE(d != nullptr, "conv3d: null descriptor");
  DAMVS_REQUIRE(d->B > 0 && d->B <= 65535 && d->Din > 0 && d->Hin > 0 && d->Win > 0, "conv3d: bad extent B=%d D=%d H=%d W=%d",
                d->B, d->Din, d->Hin, d->Win);
  DAMVS_REQUIRE(d->Cin > 0 && d->Cin % 8 == 0, "conv3d: Cin=%d must be a positive multiple of 8", d->Cin);
  if (d->plain_out)
    DAMVS_REQUIRE(d->Cout == 1, "conv3d: plain_out needs Cout == 1 (got %d)", d->Cout);
  else
    DAMVS_REQUIRE(d->Cout > 0 && d->Cout % 8 == 0, "conv3d: Cout=%d must be a positive multiple of 8", d->Cout);
  DAMVS_REQUIRE(d->transposed == 0 || d->transposed == 1, "conv3d: transposed must be 0/1");
  DAMVS_REQUIRE(d->transposed || d->stride == 1 || d->stride == 2, "conv3d: stride=%d must be 1 or 2", d->stride);
  DAMVS_REQUIRE(d->in_dtype == DAMVS_F32 || d->in_dtype == DAMVS_BF16 || d->in_dtype == DAMVS_F16, "conv3d: bad in_dtype");
  DAMVS_REQUIRE(d->out_dtype == DAMVS_F32 || d->out_dtype == DAMVS_BF16 || d->out_dtype == DAMVS_F16, "conv3d: bad out_dtype");
  DAMVS_REQUIRE(d->impl == DAMVS_CONV_DIRECT || d->impl == DAMVS_CONV_TCGEN05, "conv3d: bad impl %d", d->impl);
  return DAMVS_OK;
}

}  // namespace damvs

using namespace damvs;

extern "C" int damvs_abi_version(void) { return DAMVS_ABI_VERSION; }
extern "C" const char* damvs_last_error(void) { return g_err; }
extern "C" uint64_t damvs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int damvs_check_device(int dev) {
  cudaDeviceProp prop;
  DAMVS_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return set_error(DAMVS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                     prop.major, prop.minor);
  return DAMVS_OK;
}

extern "C" size_t damvs_conv3d_packed_weight_bytes(const damvs_conv3d_desc* d) {
  if (check_conv_desc(d) != DAMVS_OK) return 0;
  return d->impl == DAMVS_CONV_TCGEN05 ? conv3d_tc_packed_bytes(d) : conv3d_direct_packed_bytes(d);
}

extern "C" int damvs_conv3d_pack_weight(const damvs_conv3d_desc* d, const float* weight, void* packed, void* stream) {
  int rc = check_conv_desc(d);
  if (rc) return rc;
  DAMVS_REQUIRE(weight && packed, "conv3d_pack_weight: null pointer");
  DAMVS_REQUIRE(aligned16(packed), "conv3d_pack_weight: packed must be 16-byte aligned");
  return d->impl == DAMVS_CONV_TCGEN05 ? conv3d_tc_pack(d, weight, packed, (cudaStream_t)stream)
                                       : conv3d_direct_pack(d, weight, packed, (cudaStream_t)stream);
}

extern "C" int damvs_conv3d_fwd(const damvs_conv3d_desc* d, const void* in, const void* packed, const float* scale,
                                const float* shift, const void* skip, void* out, void* stream) {
  int rc = check_conv_desc(d);
  if (rc) return rc;
  DAMVS_REQUIRE(in && packed && out, "conv3d: null pointer");
  DAMVS_REQUIRE((scale == nullptr) == (shift == nullptr), "conv3d: scale and shift must both be given or both be null");
  DAMVS_REQUIRE(aligned16(in) && aligned16(out) && aligned16(packed) && aligned16(skip), "conv3d: pointers must be 16-byte aligned");
  DAMVS_REQUIRE(!(d->plain_out && skip), "conv3d: plain_out has no skip input");
  return d->impl == DAMVS_CONV_TCGEN05
             ? conv3d_tc_launch(d, in, packed, scale, shift, skip, out, (cudaStream_t)stream)
             : conv3d_direct_launch(d, in, packed, scale, shift, skip, out, (cudaStream_t)stream);
}

extern "C" int damvs_prob_head_supported(const damvs_conv3d_desc* d) {
  if (check_conv_desc(d) != DAMVS_OK) return 0;
  return d->impl == DAMVS_CONV_TCGEN05 && conv3d_tcp_head_supported(d) ? 1 : 0;
}

extern "C" int damvs_prob_head_fwd(const damvs_conv3d_desc* d, const void* in, const void* packed, const float* depth_hyp, float* prob,
                                   float* depth, float* conf, float* var, void* stream) {
  int rc = check_conv_desc(d);
  if (rc) return rc;
  DAMVS_REQUIRE(in && packed && depth_hyp && prob && depth && conf && var, "prob_head: null pointer");
  DAMVS_REQUIRE(aligned16(in) && aligned16(packed), "prob_head: in and packed must be 16-byte aligned");
  if (!(d->impl == DAMVS_CONV_TCGEN05 && conv3d_tcp_head_supported(d)))
    return set_error(DAMVS_ERR_UNSUPPORTED, "prob_head: needs the tcgen05 prob layer (Cin 8 -> 1, stride 1, plain_out, bf16 / fp16 volume) "
                                             "with D %% 8 == 0 and D <= 64; use damvs_conv3d_fwd + damvs_softmax_regress_fwd otherwise");
  return conv3d_tcp_head_launch(d, in, packed, depth_hyp, prob, depth, conf, var, (cudaStream_t)stream);
}
