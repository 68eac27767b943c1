/*
 * compactb200.h -- C ABI of libcompactb200.so, the B200 (sm_100a) implementation of the
 * CompactFusion residual-compression hot path.
 *
 * The reference has no FFI of its own: its boundary is the Python module API
 * `xfuser.compact.*`.  Each entry point below replaces the tensor
 * work of one reference function (cited as file:line under /root/reference); the Python
 * mirror (the `compactfusion_b200` modules) keeps the reference's names and signatures and binds
 * these symbols with ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Unless a function says "host", every
 *     data pointer is a DEVICE pointer on the current CUDA device.
 *   - activations are (N, C) row-major fp16 (IEEE binary16), C % 8 == 0, 16-byte aligned.
 *     Code / scale outputs may be arbitrarily (2-byte) aligned so they can point straight
 *     into the flat fp16 wire payload -- no torch.cat needed.
 *   - `stream` is a cudaStream_t (0 = legacy default stream).  Functions only enqueue work:
 *     no allocation, no synchronisation, CUDA-graph capturable.
 *   - `workspace` is caller-owned scratch of at least cf_workspace_bytes(...) bytes,
 *     256-byte aligned; contents need not be preserved between calls.
 *   - return value: 0 on success, negative cf_status on error; cf_last_error() gives text.
 *   - batched calls process `batch` (<= CF_MAX_BATCH) same-shape tensors in ONE launch per
 *     pass (K and V, or all peers of an all-gather); argument arrays are HOST arrays of
 *     device pointers, copied into kernel parameters.
 */
#ifndef COMPACTB200_H_
#define COMPACTB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CF_API __attribute__((visibility("default")))
#else
#define CF_API
#endif

#define CF_ABI_VERSION 1
#define CF_MAX_BATCH 16
#define CF_MAX_PEERS 16
#define CF_MAX_FANOUT 32 /* batch * n_dst of cf_sign_compress_put */

typedef void* cf_stream_t; /* cudaStream_t */

enum cf_status {
  CF_OK = 0,
  CF_ERR_ARG = -1,       /* bad shape / alignment / null pointer */
  CF_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  CF_ERR_WORKSPACE = -3, /* workspace too small */
  CF_ERR_UNSUPPORTED = -4
};

enum cf_codec {
  CF_CODEC_BINARY = 1, /* COMPACT_COMPRESS_TYPE.BINARY  (utils.py:21) */
  CF_CODEC_INT2 = 2,   /* COMPACT_COMPRESS_TYPE.INT2    (utils.py:22) */
  CF_CODEC_INT4 = 4,   /* COMPACT_COMPRESS_TYPE.INT4    (utils.py:24) */
  CF_CODEC_INT8 = 8,   /* quantize_int8 (compress_quantize.py:428) */
  CF_CODEC_TOPK = 16,  /* COMPACT_COMPRESS_TYPE.SPARSE  (utils.py:20) */
  CF_CODEC_LOWRANK = 32 /* COMPACT_COMPRESS_TYPE.LOW_RANK / LOW_RANK_Q (utils.py:26-27) */
};

CF_API int cf_abi_version(void);
CF_API const char* cf_last_error(void);
/* number of SMs of the current device (148 on B200); <0 on error */
CF_API int cf_sm_count(void);

/* Scratch bytes needed by the compress entry points of `codec` for `batch` tensors of
 * shape (N, C); `rank` only matters for CF_CODEC_LOWRANK. */
CF_API size_t cf_workspace_bytes(int codec, int64_t N, int64_t C, int rank, int batch);

/* ---- BINARY: 1-bit sign + rank-1 (token x channel) mean-|delta| scale ------------------
 * replaces binary_quant_fastpath + _binary_quant_fastpath (fastpath.py:124-228, :13-120)
 * with rank = -1:   delta = x - base;  V[c] = mean_n |delta|;  u[n] = mean_c |delta|;
 * U[n] = u[n] / mean_n u[n];  bit = (delta >= 0) packed LSB-first along C;
 * if new_base != NULL:  new_base = base + (2 bit - 1) * fp16(U[n] V[c])   (error feedback).
 * base == NULL means base = 0 (quantize_1bit, compress_quantize.py:7-90).
 * new_base may alias base (in-place cache update).
 * outputs: packed (N, C/8) u8, scale_u (N,1) fp16, scale_v (C,1) fp16. */
CF_API int cf_binary_compress(const void* x, const void* base, void* new_base, void* packed,
                       void* scale_u, void* scale_v, int64_t N, int64_t C, void* workspace,
                       size_t workspace_bytes, cf_stream_t stream);
CF_API int cf_binary_compress_batched(int batch, const void* const* x, const void* const* base,
                               void* const* new_base, void* const* packed,
                               void* const* scale_u, void* const* scale_v, int64_t N,
                               int64_t C, void* workspace, size_t workspace_bytes,
                               cf_stream_t stream);
/* replaces binary_dequant_fastpath + kernel (fastpath.py:371-438, :277-367):
 * recon = base + (2 bit - 1) * fp16(sum_k U[n,k] V[c,k]);  base == NULL gives the bare
 * dequantised delta (dequantize_1bit, compress_quantize.py:154-225).  K = 1 is bit-exact;
 * K > 1 (deprecated in the reference, main.py:188-189) accumulates in fp32.
 * recon may alias base.  scale_v is (C, K). */
CF_API int cf_binary_decompress(const void* packed, const void* scale_u, const void* scale_v, int K,
                         const void* base, void* recon, int64_t N, int64_t C,
                         cf_stream_t stream);
CF_API int cf_binary_decompress_batched(int batch, const void* const* packed,
                                 const void* const* scale_u, const void* const* scale_v,
                                 const void* const* base, void* const* recon, int64_t N,
                                 int64_t C, cf_stream_t stream);

/* ---- INT2: sign + magnitude bit, levels +-0.5 thr / +-2 thr, thr = fp16(chan[c] tok[n]) --
 * replaces int2_quant_fastpath + kernel (fastpath.py:584-669, :486-580) and
 * int2_dequant_fastpath + kernel (fastpath.py:745-811, :672-741).
 * outputs: packed (N, C/4) u8, scale_u = tok (N,1), scale_v = chan (C,1). */
CF_API int cf_int2_compress(const void* x, const void* base, void* new_base, void* packed,
                     void* scale_u, void* scale_v, int64_t N, int64_t C, void* workspace,
                     size_t workspace_bytes, cf_stream_t stream);
CF_API int cf_int2_compress_batched(int batch, const void* const* x, const void* const* base,
                             void* const* new_base, void* const* packed,
                             void* const* scale_u, void* const* scale_v, int64_t N, int64_t C,
                             void* workspace, size_t workspace_bytes, cf_stream_t stream);
CF_API int cf_int2_decompress(const void* packed, const void* scale_u, const void* scale_v,
                       const void* base, void* recon, int64_t N, int64_t C,
                       cf_stream_t stream);
CF_API int cf_int2_decompress_batched(int batch, const void* const* packed,
                               const void* const* scale_u, const void* const* scale_v,
                               const void* const* base, void* const* recon, int64_t N,
                               int64_t C, cf_stream_t stream);
/* Elementwise stage only, with caller-supplied scales (parity tests: codes are bit-exact
 * given identical scale tensors; SURVEY.md section 7 hard part 2). */
CF_API int cf_int2_encode_with_scales(const void* x, const void* base, const void* scale_u,
                               const void* scale_v, void* new_base, void* packed, int64_t N,
                               int64_t C, cf_stream_t stream);

/* Profiling hook: run only the selected passes of cf_{binary,int2}_compress_batched, so that
 * bench.py can time each kernel of the compress call with CUDA events on its own.
 * CF_PASS_STATS = delta statistics (+ BINARY sign bits), CF_PASS_FINALIZE = scale vectors,
 * CF_PASS_ENCODE = INT2 codes / error-feedback base.  Later passes read what earlier ones
 * left in `workspace`; CF_PASS_ALL is the ordinary compress call. */
enum cf_pass { CF_PASS_STATS = 1, CF_PASS_FINALIZE = 2, CF_PASS_ENCODE = 4, CF_PASS_ALL = 7 };
/* Optional flag OR-ed into the `codec` argument of cf_sign_compress_passes and
 * cf_sign_decompress_batched_wait.  The library's kernels are chained with programmatic
 * dependent launch; with this flag the caller promises that the tensors the call only READS
 * (x, base) were NOT written by the kernel launched immediately before it on `stream`, so
 * their first tiles are fetched before that kernel has drained.  A whole-step runtime that
 * walks distinct per-layer buffers (compactfusion_b200/engine.py) can promise this; a caller
 * that decompresses and re-compresses the same cache entry back to back cannot. */
enum cf_flag { CF_FLAG_INPUTS_STABLE = 0x100 };
CF_API int cf_sign_compress_passes(int codec, int passes, int batch, const void* const* x,
                            const void* const* base, void* const* new_base, void* const* packed,
                            void* const* scale_u, void* const* scale_v, int64_t N, int64_t C,
                            void* workspace, size_t workspace_bytes, cf_stream_t stream);

/* ---- INT4 / INT8: per-channel (over N) min/max affine codes ----------------------------
 * INT4 replaces quantize_int4 / dequantize_int4 / sim_int4(dim=0) (compress_quantize.py:
 * 487-640): scale = fp16((max-min)/(15+1e-6)), q = clamp(rne((v-min)/scale),0,15), rows
 * (2i,2i+1) share byte (i,c), low nibble = even row; N even.  v = x - base (base may be
 * NULL).  If new_base != NULL: new_base = base + (q*scale + min)  (or the bare
 * reconstruction when base == NULL: that is sim_int4).  NaN codes (zero scale) -> 0.
 * outputs: packed (N/2, C) u8, scale (1,C) fp16, minv (1,C) fp16. */
CF_API int cf_int4_compress(const void* x, const void* base, void* new_base, void* packed,
                     void* scale, void* minv, int64_t N, int64_t C, void* workspace,
                     size_t workspace_bytes, cf_stream_t stream);
CF_API int cf_int4_decompress(const void* packed, const void* scale, const void* minv,
                       const void* base, void* recon, int64_t N, int64_t C,
                       cf_stream_t stream);
/* The reference's simulation-only 4-level min/max quantiser, sim_int2_minmax (compress_quantize.py:386-426;
 * COMPACT_COMPRESS_TYPE.INT2_MINMAX in sim_compress, slowpath.py:203-204): the INT4 arithmetic with
 * qmax = 3 -- scale = fp16((max-min)/(3+1e-6)), q = clamp(rne((v-min)/scale),0,3).  Same arguments as
 * cf_int4_compress (workspace of CF_CODEC_INT4 size); the codes are written as nibbles (0..3) and are not a
 * wire format -- the result is new_base = base + (q*scale + min). */
CF_API int cf_int2mm_compress(const void* x, const void* base, void* new_base, void* packed,
                       void* scale, void* minv, int64_t N, int64_t C, void* workspace,
                       size_t workspace_bytes, cf_stream_t stream);
/* INT8 replaces quantize_int8 / dequantize_int8 (compress_quantize.py:428-484):
 * outputs: q (N,C) i8, scale (1,C) fp16, zero_point (1,C) i16. */
CF_API int cf_int8_compress(const void* x, const void* base, void* new_base, void* q, void* scale,
                     void* zero_point, int64_t N, int64_t C, void* workspace,
                     size_t workspace_bytes, cf_stream_t stream);
CF_API int cf_int8_decompress(const void* q, const void* scale, const void* zero_point,
                       const void* base, void* recon, int64_t N, int64_t C,
                       cf_stream_t stream);

/* ---- SPARSE 1:m ("top-k"): per m-block argmax |v|, lowest index wins ties ---------------
 * replaces topk_compress / topk_decompress / topk_sparsify (compress_topk.py:11-219).
 * v = x - base viewed as rows of 1024; m in {2,4,8,16}; numel % 1024 == 0.
 * outputs: val (numel/m) fp16, idx (numel/(2m)) u8 = idx_block1 << 4 | idx_block2. */
CF_API int cf_topk_compress(const void* x, const void* base, void* new_base, void* val, void* idx,
                     int64_t numel, int m, cf_stream_t stream);
CF_API int cf_topk_decompress(const void* val, const void* idx, const void* base, void* recon,
                       int64_t numel, int m, cf_stream_t stream);

/* ---- LOW_RANK: randomised subspace iteration projector ---------------------------------
 * replaces subspace_iter (compress_lowrank.py:16-62) on A = x - base (fp32 arithmetic):
 * Q <- q0 (C, r) fp32 (the caller draws / orthonormalises it, like the reference's randn+qr);
 * iters x { Z = A^T (A Q); Q = orth(Z) };  U = orth(A Q) (N, r);  V = U^T A (r, C).
 * orth() is CholeskyQR2 with an fp64 Gram matrix: it spans the same subspace as the
 * reference's Householder QR, so U V (the only comparable quantity)
 * matches to fp16 rounding.  outputs U (N,r) fp16, V (r,C) fp16. */
CF_API int cf_lowrank_project(const void* x, const void* base, const float* q0, void* U, void* V,
                       float* q_out /* (C, r) fp32 final Q, may be NULL */, int64_t N, int64_t C,
                       int rank, int iters, void* workspace, size_t workspace_bytes,
                       cf_stream_t stream);
/* replaces torch.matmul(u, v) of slowpath_decompress (slowpath.py:152-154) fused with the
 * residual add: recon = base + fp16(U V) (base may be NULL). */
CF_API int cf_lowrank_reconstruct(const void* U, const void* V, const void* base, void* recon,
                           int64_t N, int64_t C, int rank, cf_stream_t stream);
/* LOW_RANK_Q decode fused into the reconstruct: `payload` is the wire format of slowpath_compress(LOW_RANK_Q)
 * (slowpath.py:69-75): [qU (N/2, r) u8 | scaleU (r) | minU (r) | qV^T (C/2, r) u8 | scaleV (r) | minV (r)], int4 per
 * column of U and of V^T (rows 2i / 2i+1 share a byte, low nibble = even row).  recon = base + fp16(U V) with
 * U, V = the dequantised factors (value = fp16(fp16(code * scale) + min)): replaces the two dequantize_int4 calls,
 * the transpose and torch.matmul of slowpath_decompress (slowpath.py:156-164) plus the residual add.
 * recon may alias base.  N even, C % 8 == 0. */
/* LOW_RANK_Q wire packing: U (N, r) and V (r, C) fp16 -> payload in the layout above.  Replaces quantize_int4(u),
 * quantize_int4(v.t().contiguous()) and the torch.cat of slowpath_compress (slowpath.py:62-75); V^T is never
 * formed.  Bit-identical to cf_minmax_compress(CF_CODEC_INT4) on U and on V^T. */
CF_API int cf_lowrank_q_pack(const void* U, const void* V, void* payload, int64_t N, int64_t C, int rank,
                      cf_stream_t stream);
CF_API int cf_lowrank_q_reconstruct(const void* payload, const void* base, void* recon, int64_t N,
                             int64_t C, int rank, cf_stream_t stream);

/* ---- one-sided NVLink transport of the payloads (replaces dist.all_gather of
 * compact_all_gather, main.py:409, and the ring's batch_isend_irecv, ring.py:268-269) -----
 * cf_ipc_alloc: cudaMalloc + zero a buffer and export its 64-byte CUDA IPC handle (HOST
 * buffer `handle64`); cf_ipc_open maps a peer process's buffer (peer access is enabled
 * lazily); cf_ipc_close / cf_ipc_free undo them.  These four calls synchronise the device.
 * cf_p2p_put: copy `bytes` (multiple of 16) from the local payload `src` to `n_peers`
 * destinations (device pointers, local or peer-mapped), then publish ++(*local_count) to every
 * peer_flag[q] with release semantics.  local_count / local_ticket are local device u32 words
 * (ticket must start at 0).  Enqueues one kernel; CUDA-graph capturable. */
CF_API int cf_ipc_alloc(size_t bytes, void** dev_ptr, void* handle64);
CF_API int cf_ipc_open(const void* handle64, void** peer_ptr);
CF_API int cf_ipc_close(void* peer_ptr);
CF_API int cf_ipc_free(void* dev_ptr);
CF_API int cf_p2p_put(const void* src, size_t bytes, int n_peers, void* const* peer_dst,
               void* const* peer_flag, void* local_count, void* local_ticket, cf_stream_t stream);
/* cf_p2p_wait: block the STREAM (one warp on the device, no host involvement) until *flags[i] >= *expected for
 * i < n -- the counters cf_p2p_put / cf_sign_compress_put publish -- for consumers that do not wait inside their
 * own kernel (the low-rank reconstruct; the BINARY / INT2 reconstruct waits itself,
 * cf_sign_decompress_batched_wait).  A wait beyond ~2 s sets *error_word = 1 and lets the stream go on.
 * Replaces the completion wait of the ring's batch_isend_irecv (ring.py:268-269) / of dist.all_gather. */
CF_API int cf_p2p_wait(int n, const void* const* flags, const void* expected, void* error_word,
                cf_stream_t stream);
/* Fused compress + put: cf_sign_compress_passes (no cache update) whose kernels store the payload of
 * tensor t -- [codes | U (N) | V (C)], the App-A wire layout -- straight into `n_dst` receive slots
 * instead of a local send buffer: dst_payload is a HOST array of batch * n_dst device pointers, entry
 * [t * n_dst + q] = start of tensor t's payload at destination q (local memory or a peer mapping
 * from cf_ipc_open; 2-byte aligned, 16-byte aligned for the flag-waiting decompress).
 *   BINARY: the pass-1 kernel writes the sign bytes to every destination while it streams x and
 *           base; the finalize kernel writes the scale vectors and publishes the flags.
 *   INT2:   the finalize kernel writes the scale vectors; the encode kernel reads them back from
 *           destination `self_dst` (which must be LOCAL memory; ignored for BINARY), writes the code
 *           words to every destination and publishes the flags.
 * Publishing: every CTA of the call's last kernel, after a system-scope fence behind its stores,
 * adds 1 to every dst_flag[q] (an NVLink atomic for peers), and *local_count advances by that
 * kernel's grid size -- so "*flag >= *local_count" means "this put has fully landed", the
 * condition cf_sign_decompress_batched_wait waits for, whether the slot was filled by this call
 * or by cf_p2p_put (+1 per put), as long as all ranks issue the same sequence of puts.  Replaces compress + torch.cat + dist.all_gather of
 * compact_all_gather (main.py:400-409): there is no separate transport kernel or collective.
 * batch * n_dst <= CF_MAX_FANOUT; needs the pipelined kernels (base set, 64 <= C <= 8192).
 * `passes` as in cf_sign_compress_passes. */
CF_API int cf_sign_compress_put(int codec, int passes, int batch, const void* const* x,
                         const void* const* base, int n_dst, int self_dst,
                         void* const* dst_payload, void* const* dst_flag, void* local_count,
                         void* local_ticket, int64_t N, int64_t C, void* workspace,
                         size_t workspace_bytes, cf_stream_t stream);
/* cf_{binary,int2}_decompress_batched that first waits, on the device, until
 * *wait_flag[t] >= *expected for every tensor t with a non-null flag (the counters cf_p2p_put
 * publishes; `expected` is normally the caller's own local_count of the same slot).  A wait
 * that exceeds ~2 s sets *error_word = 1 instead of hanging the GPU.  Requires the pipelined
 * kernel: 16-byte aligned base/codes, C % 128 == 0 (BINARY) or C % 64 == 0 (INT2).
 * With wait_flag == expected == NULL it is the plain batched decompress (codec may carry
 * CF_FLAG_INPUTS_STABLE). */
CF_API int cf_sign_decompress_batched_wait(int codec, int batch, const void* const* packed,
                                    const void* const* scale_u, const void* const* scale_v,
                                    const void* const* base, void* const* recon,
                                    const void* const* wait_flag, const void* expected,
                                    void* error_word, int64_t N, int64_t C, cf_stream_t stream);

/* ---- consumers right behind the codec kernels ------------------------------------------
 * cf_lse_merge: the ring's per-hop online-softmax merge, replaces `update_out_and_lse` of
 * yunchang.ring.utils (third party, pinned `yunchang>=0.6.0`, setup.py:35; call sites
 * ring.py:193-195):  w = sigmoid(block_lse - lse);  out <- out - w (out - block_out);
 * lse_out <- lse_in - logsigmoid(lse_in - block_lse).
 * out (B,S,H,D) fp32, updated in place; block_out (B,S,H,D) fp16; lse_in / block_lse / lse_out
 * (B,H,S) fp32 (flash-attn's layout), lse_out must not alias lse_in.  D % 4 == 0. */
CF_API int cf_lse_merge(void* out, const void* block_out, const void* lse_in, const void* block_lse,
                 void* lse_out, int64_t B, int64_t S, int64_t H, int64_t D, cf_stream_t stream);
/* cf_error_stats: one pass over two fp16 tensors a (test) and b (reference) of `numel` elements
 * (multiple of 8, 16-byte aligned): out4 (4 x fp32, device) = { sum (a-b)^2, sum b^2, max |a-b|,
 * max |b| } -- the per-step max-abs / relative-L2 / PSNR inputs (replaces the eager torch
 * reductions of stats.py:44-120).  Deterministic (fixed reduction order, fp64 partials).
 * workspace: cf_error_stats_workspace_bytes() bytes, 256-byte aligned, ZERO-initialised once by
 * the caller (the kernel leaves it ready for the next call). */
CF_API size_t cf_error_stats_workspace_bytes(void);
CF_API int cf_error_stats(const void* a, const void* b, int64_t numel, void* out4, void* workspace,
                   size_t workspace_bytes, cf_stream_t stream);

/* ---- host-buffer entry points (end-to-end: H2D + kernels + D2H inside the call) ---------
 * x_host/base_host/recon_host are HOST (ideally pinned) buffers; payload_host receives /
 * supplies the flat fp16 wire payload [codes | U | V] (main.py:149-152).  dev_scratch is a
 * caller-owned DEVICE buffer of cf_host_scratch_bytes(codec, N, C) bytes.  The calls
 * synchronise `stream` before returning. */
CF_API size_t cf_host_scratch_bytes(int codec, int64_t N, int64_t C);
CF_API int cf_host_compress(int codec, const void* x_host, const void* base_host, void* new_base_host,
                     void* payload_host, int64_t N, int64_t C, void* dev_scratch,
                     size_t dev_scratch_bytes, cf_stream_t stream);
CF_API int cf_host_decompress(int codec, const void* payload_host, const void* base_host,
                       void* recon_host, int64_t N, int64_t C, void* dev_scratch,
                       size_t dev_scratch_bytes, cf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COMPACTB200_H_ */
