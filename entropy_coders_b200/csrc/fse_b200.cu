// fse_b200.cu -- C ABI of libfse_b200.so (see include/fse_b200.h).  Host orchestration only;
// all arithmetic runs in the kernels of fse_kernels.cuh.  There is no CPU fallback.
#include "../../include/fse_b200.h"
#include "fse_encode128.cuh"
#include "fse_decode128c.cuh"
#include "fse_hist16.cuh"
#include "fse_shared_enc.cuh"
#include "fse_shared_dec.cuh"
#include "fse_tps.cuh"
#include "fse_bitio.cuh"
#include "fse_zstd_norm.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace fsed;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = std::max(bytes, cap + cap / 2);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

struct fse_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;
    uint64_t launches = 0;
    std::string err;
    // workspaces
    DevBuf counts, hist_pieces, hlen, plen, scratch, offsets_tmp, status_tmp, misc;
    DevBuf tps_enc_tab, tps_enc_tt, tps_dec_tab, tps_meta;      // per-block tables of the thread-per-stream coders (fse_tps.cuh)
    DevBuf stage_in, stage_out, stage_off, stage_status;  // host-buffer conveniences
    HostBuf pin;
    // global table
    DevBuf g_enc_table, g_enc_tt, g_dec_table, g_norm, g_meta, g_hdr;
    uint32_t g_log2 = 0, g_table_len = 0;
    bool g_valid = false;
    // optional per-kernel timing (bench.py's roofline figures): events on the launching stream
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int kernel; size_t e0, e1; };
    std::vector<Span> spans;
    double t_ms[FSE_B200_NUM_KERNELS] = {0};
    uint64_t t_count[FSE_B200_NUM_KERNELS] = {0};
    // copy streams of the pipelined host entry points
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> pipe_ev;
    // generator LUTs
    DevBuf lut[4];
    uint32_t lut_len[4] = {0, 0, 0, 0};
};

namespace {

int fail(fse_b200_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call)                                                             \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return fail(ctx, FSE_B200_ERR_CUDA, #call, e_); \
    } while (0)

// RAII-less span helper: records an event pair around one kernel launch when timing is on
size_t ev_get(fse_b200_ctx *c)
{
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_used++;
}
struct Timed {
    fse_b200_ctx *c; int k; size_t e0 = 0;
    Timed(fse_b200_ctx *ctx, int kernel) : c(ctx), k(kernel)
    {
        if (c->timing) { e0 = ev_get(c); cudaEventRecord(c->ev_pool[e0], c->stream); }
    }
    ~Timed()
    {
        c->launches++;
        if (c->timing) {
            size_t e1 = ev_get(c);
            cudaEventRecord(c->ev_pool[e1], c->stream);
            c->spans.push_back({k, e0, e1});
        }
    }
};

bool pow2(uint32_t v) { return v && !(v & (v - 1)); }

// Streams per CTA of the shared-memory thread-per-stream kernels (one CTA per SM at a time): `fit` table sets fit in shared
// memory; the streams are spread evenly over the fewest rounds of num_sms CTAs, so that the last round is as full as the
// others (8 192 blocks, 37 sets fit: 2 rounds of 28 instead of one of 37 and half a round)
uint32_t tps_streams_per_cta(size_t streams, size_t num_sms, size_t fit)
{
    if (fit < 1) return 0;
    const size_t rounds = (streams + num_sms * fit - 1) / (num_sms * fit);
    const size_t per = (streams + num_sms * rounds - 1) / (num_sms * std::max<size_t>(rounds, 1));
    return (uint32_t)std::max<size_t>(1, std::min(per, fit));
}

// development overrides (variant builds under tools/bin only): the release library ignores the environment
int dev_opt(const char *name, int dflt)
{
#ifdef FSE_DEV
    if (const char *o = getenv(name)) return atoi(o);
#else
    (void)name;
#endif
    return dflt;
}

int check_params(fse_b200_ctx *ctx, const fse_b200_params *p)
{
    if (!ctx || !p) return FSE_B200_ERR_ARG;
    if (p->block_size == 0 || p->block_size > (1u << 30)) return fail(ctx, FSE_B200_ERR_ARG, "block_size must be in 1..2^30");
    if (!pow2(p->n_states) || p->n_states > 128) return fail(ctx, FSE_B200_ERR_ARG, "n_states must be a power of two up to 128");
    if (p->n_states >= 64 && p->table_log > 13) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "n_states 64 / 128 need table_log <= 13");
    if (p->table_log != 0 && (p->table_log < 5 || p->table_log > 15)) return fail(ctx, FSE_B200_ERR_ARG, "table_log must be 0 or 5..15");
    if (p->table_mode > 1) return fail(ctx, FSE_B200_ERR_ARG, "table_mode");
    if (p->segment_size) {
        if (p->table_mode != FSE_B200_TABLE_PER_BLOCK || p->n_states != 128)
            return fail(ctx, FSE_B200_ERR_ARG, "segment_size needs per-block tables and n_states 128");
        if (p->segment_size < 512 || p->block_size % p->segment_size || p->block_size / p->segment_size > 64)
            return fail(ctx, FSE_B200_ERR_ARG, "segment_size must be >= 512 and divide block_size into at most 64 segments");
        if (p->table_log > SH_TL_MAX) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "segment_size needs table_log <= 11");
    }
    if (p->flags & ~FSE_B200_FLAG_RAW_IF_EXPANDS) return fail(ctx, FSE_B200_ERR_ARG, "unknown flags");
    if (p->flags && (p->segment_size || p->table_mode != FSE_B200_TABLE_PER_BLOCK))
        return fail(ctx, FSE_B200_ERR_ARG, "FSE_B200_FLAG_RAW_IF_EXPANDS needs per-block tables and one stream per block");
    return 0;
}

// entries of the stream index and the bytes one stream covers
size_t stream_bytes(const fse_b200_params *p) { return p->segment_size ? p->segment_size : p->block_size; }
size_t num_streams(size_t n, const fse_b200_params *p)
{
    const size_t u = stream_bytes(p);
    return u ? (n + u - 1) / u : 0;
}

// largest table_log a per-block launch can meet: the request (or optimal_log2 <= 11), raised to
// ilog2(table_len-1)+2 <= 9 (src/histogram.rs:96-98)
uint32_t tlmax_for(const fse_b200_params *p) { return p->table_log == 0 ? 11u : std::max(p->table_log, 9u); }

// Payload capacity of one stream.  A stream coded with the table of ITS OWN bytes cannot grow past compress_bound
// (src/fse.rs:191-193).  A stream coded with a table made from other bytes as well (global table, segments of a block)
// can cost up to table_log bits per symbol, whatever it holds: the slot is sized for that.
size_t pay_cap_bytes(const fse_b200_params *p)
{
    const size_t sb = stream_bytes(p), ns = p->n_states ? p->n_states : 32;
    size_t v = sb + (sb >> 7) + 2 * ns + 64;
    if (p->table_mode == FSE_B200_TABLE_GLOBAL || p->segment_size) {
        const size_t tl = p->table_log ? std::max<size_t>(p->table_log, 10) : 11;   // the effective table_log may be raised to 10 (histogram.rs:96-98)
        v = std::max(v, (sb * tl + 7) / 8 + 2 * ns + 64);
    }
    return (v + 15) & ~(size_t)15;
}
size_t scratch_stride(const fse_b200_params *p) { return HDR_RESERVE + pay_cap_bytes(p); }

// pick warps per CTA so that the last wave of blocks is as full as possible
int pick_warps(size_t nblocks, int num_sms, size_t per_warp_smem, size_t smem_limit, int max_warps)
{
    int wmax = (int)std::min<size_t>((size_t)max_warps, smem_limit / per_warp_smem);
    if (wmax < 1) return 0;
    int best = wmax;
    double best_eff = -1;
    for (int w = wmax; w >= std::max(1, (wmax * 3) / 4); w--) {
        size_t slots = (size_t)num_sms * w;
        size_t waves = (nblocks + slots - 1) / slots;
        double eff = (double)nblocks / (double)(waves * slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = w; }
    }
    return best;
}

// Histogram::new per block: 16-bit lane-private columns, one warp per block (blocks up to 1 MiB).  When the blocks are too
// few to fill the warp slots of the machine (or larger than 1 MiB) each block is counted as k equal pieces, one warp per
// piece, and the pieces are summed (integer sums: the result is the same).  Blocks that do not split evenly keep the
// one-CTA-per-block kernel with 32-bit columns.
int launch_hist(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint32_t block_size, size_t nb, uint32_t *d_counts,
                uint32_t *d_table_len)
{
    const size_t slots = (size_t)ctx->num_sms * HIST16_WARPS;
    uint32_t k = 1;
    if (nb < slots || block_size > HIST16_MAX_BLOCK)
        while ((nb * k < 2 * slots || block_size / k > HIST16_MAX_BLOCK) && block_size % (2 * k) == 0 && (block_size / (2 * k)) % 16 == 0 &&
               block_size / (2 * k) >= 8192)
            k *= 2;
    if (k > 1) {
        const uint32_t piece = block_size / k;
        const size_t npieces = (n + piece - 1) / piece;
        if (npieces <= 0xffffffffull) {
            CK(ctx->hist_pieces.reserve(npieces * 256 * sizeof(uint32_t)));
            int grid = (int)std::min<size_t>((npieces + HIST16_WARPS - 1) / HIST16_WARPS, (size_t)ctx->num_sms);
            k_hist_blocks16<<<grid, HIST16_WARPS * 32, HIST16_SMEM, ctx->stream>>>(d_src, n, piece, (uint32_t)npieces,
                                                                                   ctx->hist_pieces.as<uint32_t>(), nullptr);
            k_hist_sum_pieces<<<(int)std::min<size_t>(nb, (size_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(
                ctx->hist_pieces.as<uint32_t>(), (uint32_t)npieces, k, (uint32_t)nb, d_counts, d_table_len);
            ctx->launches++;
            return FSE_B200_OK;
        }
    }
    if (block_size <= HIST16_MAX_BLOCK) {
        int grid = (int)std::min<size_t>((nb + HIST16_WARPS - 1) / HIST16_WARPS, (size_t)ctx->num_sms);
        k_hist_blocks16<<<grid, HIST16_WARPS * 32, HIST16_SMEM, ctx->stream>>>(d_src, n, block_size, (uint32_t)nb, d_counts, d_table_len);
    } else {
        int grid = (int)std::min<size_t>(nb, (size_t)ctx->num_sms * 3);
        k_hist_blocks<<<grid, HIST_WARPS * 32, HIST_SMEM, ctx->stream>>>(d_src, n, block_size, (uint32_t)nb, d_counts, d_table_len);
    }
    return FSE_B200_OK;
}

}  // namespace

extern "C" {

const char *fse_b200_version(void) { return "fse_b200 0.1 (sm_100a)"; }

int fse_b200_create(int device, void *stream, fse_b200_ctx **out)
{
    if (!out) return FSE_B200_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return FSE_B200_ERR_CUDA;
    fse_b200_ctx *ctx = new (std::nothrow) fse_b200_ctx();
    if (!ctx) return FSE_B200_ERR_CUDA;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return FSE_B200_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return FSE_B200_ERR_CUDA; }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    if (stream) ctx->stream = (cudaStream_t)stream;
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FSE_B200_ERR_CUDA; }
        ctx->own_stream = true;
    }
    cudaFuncSetAttribute(k_hist_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_SMEM);
    cudaFuncSetAttribute(k_hist_blocks16, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST16_SMEM);
    cudaFuncSetAttribute(k_encode_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_decode_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_encode64_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_decode64_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_decode64c_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_decode128_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);   // 16 bytes of static shared memory (block queue)
    cudaFuncSetAttribute(k_decode128c_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);   // 16 bytes of static shared memory (block queue)
    cudaFuncSetAttribute(k_encode128_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);   // 16 bytes of static shared memory (block queue)
    cudaFuncSetAttribute(k_build_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_tps_prepare_enc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_tps_prepare_dec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_tps_decode_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_tps_decode_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_tps_encode_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin);
    cudaFuncSetAttribute(k_encode_sh_global<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);
    cudaFuncSetAttribute(k_decode_sh_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);
    cudaFuncSetAttribute(k_encode_sh_blocks<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);
    cudaFuncSetAttribute(k_decode_sh_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fse_b200_destroy(ctx); return FSE_B200_ERR_CUDA; }
    *out = ctx;
    return FSE_B200_OK;
}

void fse_b200_destroy(fse_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->tps_enc_tab, &ctx->tps_enc_tt, &ctx->tps_dec_tab, &ctx->tps_meta, &ctx->counts, &ctx->hist_pieces, &ctx->hlen, &ctx->plen, &ctx->scratch, &ctx->offsets_tmp, &ctx->status_tmp, &ctx->misc,
                      &ctx->stage_in, &ctx->stage_out, &ctx->stage_off, &ctx->stage_status, &ctx->g_enc_table,
                      &ctx->g_enc_tt, &ctx->g_dec_table, &ctx->g_norm, &ctx->g_meta, &ctx->g_hdr,
                      &ctx->lut[0], &ctx->lut[1], &ctx->lut[2], &ctx->lut[3]};
    for (DevBuf *b : bufs) b->release();
    ctx->pin.release();
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->pipe_ev) cudaEventDestroy(e);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *fse_b200_last_error(const fse_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t fse_b200_launch_count(const fse_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fse_b200_set_timing(fse_b200_ctx *ctx, int enable)
{
    if (!ctx) return FSE_B200_ERR_ARG;
    ctx->timing = enable != 0;
    ctx->spans.clear();
    ctx->ev_used = 0;
    for (int k = 0; k < FSE_B200_NUM_KERNELS; k++) { ctx->t_ms[k] = 0; ctx->t_count[k] = 0; }
    return FSE_B200_OK;
}

int fse_b200_get_timing(fse_b200_ctx *ctx, double *ms_total, uint64_t *count)
{
    if (!ctx || !ms_total || !count) return FSE_B200_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (const auto &sp : ctx->spans) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev_pool[sp.e0], ctx->ev_pool[sp.e1]));
        ctx->t_ms[sp.kernel] += ms;
        ctx->t_count[sp.kernel] += 1;
    }
    ctx->spans.clear();
    ctx->ev_used = 0;
    for (int k = 0; k < FSE_B200_NUM_KERNELS; k++) { ms_total[k] = ctx->t_ms[k]; count[k] = ctx->t_count[k]; }
    return FSE_B200_OK;
}

int fse_b200_sync(fse_b200_ctx *ctx)
{
    if (!ctx) return FSE_B200_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

size_t fse_b200_compress_bound(size_t size) { return 512 + size + (size >> 7) + 4 + 8; }
size_t fse_b200_num_blocks(size_t n, uint32_t block_size) { return block_size ? (n + block_size - 1) / block_size : 0; }
size_t fse_b200_num_streams(size_t n, const fse_b200_params *p) { return p && p->block_size ? num_streams(n, p) : 0; }
size_t fse_b200_compress_blocks_bound(size_t n, const fse_b200_params *p)
{
    if (!p || !p->block_size) return 0;
    return num_streams(n, p) * scratch_stride(p) + 16;
}

// ---------------------------------------------------------------------------------- stages

int fse_b200_histogram_blocks(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint32_t block_size,
                              uint32_t *d_counts, uint32_t *d_table_len)
{
    if (!ctx || !d_src || !d_counts || block_size == 0) return fail(ctx, FSE_B200_ERR_ARG, "histogram_blocks: bad argument");
    CK(cudaSetDevice(ctx->device));
    size_t nb = fse_b200_num_blocks(n, block_size);
    if (nb == 0) return FSE_B200_OK;
    if (nb > 0xffffffffull) return fail(ctx, FSE_B200_ERR_ARG, "too many blocks");
    { int hrc = launch_hist(ctx, d_src, n, block_size, nb, d_counts, d_table_len); if (hrc) return hrc; }
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

static int hist_global_async(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint64_t *d_counts64)
{
    // Pieces of at most 1 MiB (their counts fit the 16-bit lane columns); sized so that the pieces fill four waves of
    // the histogram kernel's warp slots.  Integer sums: the result does not depend on the piece size.
    const size_t slots = (size_t)ctx->num_sms * HIST16_WARPS * 4;
    size_t want = ((n + slots - 1) / slots + 15) & ~(size_t)15;
    const uint32_t piece = (uint32_t)std::min<size_t>(std::max<size_t>(want, 32u << 10), 1u << 20);
    size_t nb = fse_b200_num_blocks(n, piece);
    CK(cudaMemsetAsync(d_counts64, 0, 256 * sizeof(uint64_t), ctx->stream));
    if (nb == 0) return FSE_B200_OK;
    CK(ctx->counts.reserve(nb * 256 * sizeof(uint32_t)));
    { int hrc = launch_hist(ctx, d_src, n, piece, nb, ctx->counts.as<uint32_t>(), nullptr); if (hrc) return hrc; }
    ctx->launches++;
    k_hist_reduce<<<(int)std::min<size_t>(nb, (size_t)ctx->num_sms * 2), 256, 0, ctx->stream>>>(ctx->counts.as<uint32_t>(), (uint32_t)nb,
                                                                          reinterpret_cast<unsigned long long *>(d_counts64));
    ctx->launches++;
    CK(cudaGetLastError());
    return FSE_B200_OK;
}

int fse_b200_histogram_global(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint64_t *d_counts64)
{
    if (!ctx || (!d_src && n) || !d_counts64) return fail(ctx, FSE_B200_ERR_ARG, "histogram_global: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rc = hist_global_async(ctx, d_src, n, d_counts64);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_normalize(fse_b200_ctx *ctx, const uint64_t *d_counts64, size_t ntables, uint32_t table_log,
                       int32_t *d_norm, uint32_t *d_log2, uint32_t *d_table_len, int32_t *d_status)
{
    if (!ctx || !d_counts64 || !d_norm || !d_log2 || !d_table_len || !d_status) return fail(ctx, FSE_B200_ERR_ARG, "normalize: bad argument");
    if (table_log != 0 && (table_log < 5 || table_log > 15)) return fail(ctx, FSE_B200_ERR_ARG, "table_log must be 0 or 5..15");
    CK(cudaSetDevice(ctx->device));
    if (ntables == 0) return FSE_B200_OK;
    int grid = (int)std::min<size_t>(ntables, (size_t)ctx->num_sms * 16);
    k_normalize<<<grid, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(d_counts64), (uint32_t)ntables,
                                              table_log, d_norm, d_log2, d_table_len, d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_normalize_zstd(fse_b200_ctx *ctx, const uint64_t *d_counts64, size_t ntables, uint32_t table_log, int use_low_prob_count,
                            int32_t *d_norm, uint32_t *d_log2, uint32_t *d_table_len, int32_t *d_status)
{
    if (!ctx || !d_counts64 || !d_norm || !d_log2 || !d_table_len || !d_status) return fail(ctx, FSE_B200_ERR_ARG, "normalize_zstd: bad argument");
    if (table_log > 15) return fail(ctx, FSE_B200_ERR_ARG, "table_log must be 0..15 (values the algorithm rejects are reported per table)");
    CK(cudaSetDevice(ctx->device));
    if (ntables == 0) return FSE_B200_OK;
    k_normalize_zstd<<<(unsigned)((ntables + 63) / 64), 64, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(d_counts64), (uint32_t)ntables,
                                                                              table_log, use_low_prob_count, d_norm, d_log2, d_table_len, d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_ncount_write(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2, const uint32_t *d_table_len,
                          size_t ntables, uint8_t *d_out, size_t stride, uint32_t *d_bytes, uint32_t *d_bits)
{
    if (!ctx || !d_norm || !d_log2 || !d_table_len || !d_out || !d_bytes || !d_bits) return fail(ctx, FSE_B200_ERR_ARG, "ncount_write: bad argument");
    if (stride < 512) return fail(ctx, FSE_B200_ERR_CAPACITY, "ncount_write: stride must be >= 512");
    CK(cudaSetDevice(ctx->device));
    if (ntables == 0) return FSE_B200_OK;
    int grid = (int)std::min<size_t>(ntables, (size_t)ctx->num_sms * 16);
    k_ncount_write<<<grid, 32, 0, ctx->stream>>>(d_norm, d_log2, d_table_len, (uint32_t)ntables, d_out, stride, d_bytes, d_bits);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_ncount_read(fse_b200_ctx *ctx, const uint8_t *d_in, size_t stride, const uint32_t *d_len, size_t ntables,
                         int32_t *d_norm, uint32_t *d_log2, uint32_t *d_table_len, uint32_t *d_consumed, int32_t *d_status)
{
    if (!ctx || !d_in || !d_len || !d_norm || !d_log2 || !d_table_len || !d_consumed || !d_status)
        return fail(ctx, FSE_B200_ERR_ARG, "ncount_read: bad argument");
    CK(cudaSetDevice(ctx->device));
    if (ntables == 0) return FSE_B200_OK;
    int grid = (int)std::min<size_t>(ntables, (size_t)ctx->num_sms * 16);
    k_ncount_read<<<grid, 32, 0, ctx->stream>>>(d_in, stride, d_len, (uint32_t)ntables, d_norm, d_log2, d_table_len, d_consumed, d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

static int build_tables(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2, const uint32_t *d_table_len,
                        size_t ntables, uint32_t max_log2, int decode, uint16_t *enc, uint2 *tt, uint8_t *sym,
                        uint32_t *dec, int32_t *d_status, bool sync)
{
    if (max_log2 < 5 || max_log2 > 15) return fail(ctx, FSE_B200_ERR_ARG, "max_table_log must be 5..15");
    CK(cudaSetDevice(ctx->device));
    if (ntables == 0) return FSE_B200_OK;
    size_t smem = 4096 + 5 * ((size_t)1 << max_log2);
    if (smem > ctx->smem_optin) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table too large for shared memory");
    int grid = (int)std::min<size_t>(ntables, (size_t)ctx->num_sms * 8);
    k_build_tables<<<grid, 32, smem, ctx->stream>>>(d_norm, d_log2, d_table_len, (uint32_t)ntables, max_log2, decode, enc, tt, sym, dec, d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    if (sync) CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_build_encode_tables(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2, const uint32_t *d_table_len,
                                 size_t ntables, uint32_t max_table_log, uint16_t *d_table,
                                 fse_b200_symbol_transform *d_symbol_tt, uint8_t *d_symbols, int32_t *d_status)
{
    if (!ctx || !d_norm || !d_log2 || !d_table_len || !d_table || !d_symbol_tt || !d_status)
        return fail(ctx, FSE_B200_ERR_ARG, "build_encode_tables: bad argument");
    return build_tables(ctx, d_norm, d_log2, d_table_len, ntables, max_table_log, 0, d_table,
                        reinterpret_cast<uint2 *>(d_symbol_tt), d_symbols, nullptr, d_status, true);
}

int fse_b200_build_decode_tables(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2, const uint32_t *d_table_len,
                                 size_t ntables, uint32_t max_table_log, fse_b200_decode_transform *d_table, int32_t *d_status)
{
    if (!ctx || !d_norm || !d_log2 || !d_table_len || !d_table || !d_status)
        return fail(ctx, FSE_B200_ERR_ARG, "build_decode_tables: bad argument");
    return build_tables(ctx, d_norm, d_log2, d_table_len, ntables, max_table_log, 1, nullptr, nullptr, nullptr,
                        reinterpret_cast<uint32_t *>(d_table), d_status, true);
}

// ---------------------------------------------------------------------------------- bit I/O primitives

int fse_b200_bitstack_write(fse_b200_ctx *ctx, const uint32_t *d_vals, const uint8_t *d_bits, size_t n, int mark,
                            uint8_t *d_out, size_t out_cap, uint64_t *h_nbits)
{
    if (!ctx || (!d_vals && n) || (!d_bits && n) || !d_out || !h_nbits || n > 0x7fffffffull) return fail(ctx, FSE_B200_ERR_ARG, "bitstack_write: bad argument");
    if (((uintptr_t)d_out & 3) != 0) return fail(ctx, FSE_B200_ERR_ARG, "bitstack_write: d_out must be 4-byte aligned");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->misc.reserve(64));
    unsigned long long *d_n = ctx->misc.as<unsigned long long>();
    int *d_st = reinterpret_cast<int *>(d_n + 1);
    k_bitstack_write<<<1, 32, 0, ctx->stream>>>(d_vals, d_bits, (uint32_t)n, mark, reinterpret_cast<uint32_t *>(d_out), (uint32_t)(out_cap / 4), d_n, d_st);
    ctx->launches++;
    CK(cudaGetLastError());
    struct { unsigned long long nbits; int st; } r;
    CK(cudaMemcpyAsync(&r, d_n, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *h_nbits = r.nbits;
    return r.st < 0 ? fail(ctx, r.st, "bitstack_write: out_cap too small") : FSE_B200_OK;
}

int fse_b200_bitstack_read(fse_b200_ctx *ctx, const uint8_t *d_in, size_t nbytes, const uint8_t *d_bits, size_t n, uint32_t *d_vals,
                           int32_t *h_status)
{
    if (!ctx || (!d_in && nbytes) || (!d_bits && n) || (!d_vals && n) || !h_status || n > 0x7fffffffull) return fail(ctx, FSE_B200_ERR_ARG, "bitstack_read: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->misc.reserve(64));
    int *d_st = ctx->misc.as<int>();
    k_bitstack_read<<<1, 32, 0, ctx->stream>>>(d_in, nbytes, d_bits, (uint32_t)n, d_vals, d_st);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_status, d_st, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_bitstream_read(fse_b200_ctx *ctx, const uint8_t *d_in, size_t nbytes, uint64_t total_bits, const uint8_t *d_bits, size_t n,
                            uint32_t *d_vals, int32_t *h_status)
{
    if (!ctx || (!d_in && nbytes) || (!d_bits && n) || (!d_vals && n) || !h_status || n > 0x7fffffffull || nbytes > 0x1fffffffull)
        return fail(ctx, FSE_B200_ERR_ARG, "bitstream_read: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->misc.reserve(64));
    int *d_st = ctx->misc.as<int>();
    k_bitstream_read<<<1, 32, 0, ctx->stream>>>(d_in, nbytes, total_bits, d_bits, (uint32_t)n, d_vals, d_st);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_status, d_st, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

// ---------------------------------------------------------------------------------- global table

// meta = the first words of g_meta read back by the caller (log2, table_len, status): one synchronisation per set-up
static int install_global(fse_b200_ctx *ctx, const uint32_t *meta, uint32_t *h_log2)
{
    // g_norm is on the device; build both tables at the effective log2 (no further synchronisation)
    int32_t st = (int32_t)meta[2];
    if (st < 0) return fail(ctx, st, "global table: normalisation failed");
    uint32_t log2 = meta[0];
    size_t size = (size_t)1 << log2;
    CK(ctx->g_enc_table.reserve(size * 2));
    CK(ctx->g_enc_tt.reserve(256 * 8));
    CK(ctx->g_dec_table.reserve(size * 4));
    uint32_t *m = ctx->g_meta.as<uint32_t>();
    int rc = build_tables(ctx, ctx->g_norm.as<int32_t>(), m, m + 1, 1, log2, 0, ctx->g_enc_table.as<uint16_t>(),
                          ctx->g_enc_tt.as<uint2>(), nullptr, nullptr, reinterpret_cast<int32_t *>(m + 3), false);
    if (rc) return rc;
    k_sanitize_global_tt<<<1, 256, 0, ctx->stream>>>(ctx->g_enc_tt.as<uint2>(), ctx->g_norm.as<int32_t>(), log2);
    ctx->launches++;
    rc = build_tables(ctx, ctx->g_norm.as<int32_t>(), m, m + 1, 1, log2, 1, nullptr, nullptr, nullptr,
                      ctx->g_dec_table.as<uint32_t>(), reinterpret_cast<int32_t *>(m + 3), false);
    if (rc) return rc;
    ctx->g_log2 = log2;
    ctx->g_table_len = meta[1];
    ctx->g_valid = true;
    if (h_log2) *h_log2 = log2;
    return FSE_B200_OK;
}

int fse_b200_set_global_table(fse_b200_ctx *ctx, const uint64_t *d_counts64, uint32_t table_log,
                              uint8_t *h_header, size_t *h_header_bytes, uint32_t *h_log2)
{
    if (!ctx || !d_counts64) return fail(ctx, FSE_B200_ERR_ARG, "set_global_table: bad argument");
    if (table_log != 0 && (table_log < 5 || table_log > 15)) return fail(ctx, FSE_B200_ERR_ARG, "table_log must be 0 or 5..15");
    CK(cudaSetDevice(ctx->device));
    ctx->g_valid = false;
    CK(ctx->g_norm.reserve(1024));
    CK(ctx->g_meta.reserve(64));
    CK(ctx->g_hdr.reserve(512));
    uint32_t *m = ctx->g_meta.as<uint32_t>();
    k_normalize<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(d_counts64), 1, table_log,
                                           ctx->g_norm.as<int32_t>(), m, m + 1, reinterpret_cast<int32_t *>(m + 2));
    ctx->launches++;
    CK(cudaGetLastError());
    const bool want_header = h_header && h_header_bytes;
    uint32_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t hdr[512];
    if (want_header) {   // the header only needs the normalised counts: written before the one read-back below
        k_ncount_write<<<1, 32, 0, ctx->stream>>>(ctx->g_norm.as<int32_t>(), m, m + 1, 1, ctx->g_hdr.as<uint8_t>(), 512, m + 4, m + 5);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hdr, ctx->g_hdr.p, sizeof(hdr), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaMemcpyAsync(meta, m, sizeof(meta), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = install_global(ctx, meta, h_log2);
    if (rc) return rc;
    if (want_header) {
        if (meta[4] > *h_header_bytes || meta[4] > sizeof(hdr)) return fail(ctx, FSE_B200_ERR_CAPACITY, "header buffer too small");
        memcpy(h_header, hdr, meta[4]);
        *h_header_bytes = meta[4];
    }
    return FSE_B200_OK;
}

int fse_b200_set_global_table_from_header(fse_b200_ctx *ctx, const uint8_t *h_header, size_t header_bytes, uint32_t *h_log2)
{
    if (!ctx || !h_header || header_bytes == 0 || header_bytes > 512) return fail(ctx, FSE_B200_ERR_ARG, "set_global_table_from_header: bad argument");
    CK(cudaSetDevice(ctx->device));
    ctx->g_valid = false;
    CK(ctx->g_norm.reserve(1024));
    CK(ctx->g_meta.reserve(64));
    CK(ctx->g_hdr.reserve(512));
    uint32_t *m = ctx->g_meta.as<uint32_t>();
    uint32_t len = (uint32_t)header_bytes;
    CK(cudaMemcpyAsync(ctx->g_hdr.p, h_header, header_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(m + 6, &len, 4, cudaMemcpyHostToDevice, ctx->stream));
    k_ncount_read<<<1, 32, 0, ctx->stream>>>(ctx->g_hdr.as<uint8_t>(), 512, m + 6, 1, ctx->g_norm.as<int32_t>(), m, m + 1, m + 7,
                                             reinterpret_cast<int32_t *>(m + 2));
    ctx->launches++;
    CK(cudaGetLastError());
    uint32_t meta[4];
    CK(cudaMemcpyAsync(meta, m, sizeof(meta), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return install_global(ctx, meta, h_log2);
}

int fse_b200_global_table_covers(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint64_t *h_unknown)
{
    if (!ctx || (!d_src && n) || !h_unknown) return fail(ctx, FSE_B200_ERR_ARG, "global_table_covers: bad argument");
    if (!ctx->g_valid) return fail(ctx, FSE_B200_ERR_ARG, "global table not installed");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->misc.reserve(256 * sizeof(uint64_t)));
    int rc = hist_global_async(ctx, d_src, n, ctx->misc.as<uint64_t>());
    if (rc) return rc;
    CK(ctx->pin.reserve(256 * sizeof(uint64_t) + 256 * sizeof(int32_t)));
    uint64_t *h_counts = reinterpret_cast<uint64_t *>(ctx->pin.p);
    int32_t *h_norm = reinterpret_cast<int32_t *>(h_counts + 256);
    CK(cudaMemcpyAsync(h_counts, ctx->misc.p, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_norm, ctx->g_norm.p, 256 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint64_t unknown = 0;
    for (uint32_t s = 0; s < 256; s++)
        if (s >= ctx->g_table_len || h_norm[s] == 0) unknown += h_counts[s];
    *h_unknown = unknown;
    return FSE_B200_OK;
}

// ---------------------------------------------------------------------------------- pipelines

int fse_b200_compress_blocks_async(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, const fse_b200_params *p,
                                   uint8_t *d_dst, size_t dst_cap, uint64_t *d_offsets, int32_t *d_status)
{
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if ((!d_src && n) || !d_dst || !d_offsets || !d_status) return fail(ctx, FSE_B200_ERR_ARG, "compress_blocks: null pointer");
    CK(cudaSetDevice(ctx->device));
    const size_t nb = fse_b200_num_blocks(n, p->block_size);      // tables (histograms) in per-block mode
    const size_t ns = num_streams(n, p);                          // entries of the index
    if (ns > 0x7fffffffull) return fail(ctx, FSE_B200_ERR_ARG, "too many blocks");
    if (dst_cap < fse_b200_compress_blocks_bound(n, p)) return fail(ctx, FSE_B200_ERR_CAPACITY, "dst_cap < fse_b200_compress_blocks_bound");
    const bool global = p->table_mode == FSE_B200_TABLE_GLOBAL;
    if (global && !ctx->g_valid) return fail(ctx, FSE_B200_ERR_ARG, "global table not installed");
    if (nb == 0) {
        CK(cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), ctx->stream));
        return FSE_B200_OK;
    }
    const uint32_t tlmax = global ? ctx->g_log2 : tlmax_for(p);
    if (global && ctx->g_log2 > (p->table_log ? std::max(p->table_log, 10u) : 11u))
        return fail(ctx, FSE_B200_ERR_ARG, "global table: pass the table_log the table was installed with (it sizes the block slots)");
    const size_t stride = scratch_stride(p);
    CK(ctx->hlen.reserve(ns * 4));
    CK(ctx->plen.reserve(ns * 4));
    CK(ctx->scratch.reserve(ns * stride));
    if (!global) {
        CK(ctx->counts.reserve(nb * 256 * sizeof(uint32_t)));
        Timed t(ctx, FSE_B200_K_HIST);
        { int hrc = launch_hist(ctx, d_src, n, p->block_size, nb, ctx->counts.as<uint32_t>(), nullptr); if (hrc) return hrc; }
    }
    EncArgs a;
    a.src = d_src; a.n = n; a.block_size = p->block_size; a.nblocks = (uint32_t)nb;
    a.req_log2 = p->table_log; a.n_states = p->n_states; a.tlmax = tlmax;
    a.counts = ctx->counts.as<uint32_t>();
    a.scratch = ctx->scratch.as<uint8_t>(); a.stride = stride;
    a.pay_cap_words = (uint32_t)(pay_cap_bytes(p) / 4);
    a.seg_size = p->segment_size; a.segs_per_block = p->segment_size ? p->block_size / p->segment_size : 1;
    a.flags = p->flags;
    a.hlen = ctx->hlen.as<uint32_t>(); a.plen = ctx->plen.as<uint32_t>(); a.status = d_status;
    a.global_mode = global ? 1 : 0;
    a.g.log2 = ctx->g_log2; a.g.table_len = ctx->g_table_len;
    a.g.enc_table = ctx->g_enc_table.as<uint16_t>(); a.g.enc_tt = ctx->g_enc_tt.as<uint2>(); a.g.dec_table = ctx->g_dec_table.as<uint32_t>();
    const bool wide = p->n_states >= 64;
    if (wide && tlmax > 13) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "n_states 64 / 128 need table_log <= 13");
    const size_t per_warp = wide ? enc64_layout(tlmax).total : enc_layout(tlmax).total;
    int wpc = pick_warps(nb, ctx->num_sms, per_warp, ctx->smem_optin, 16);
    if (p->n_states == 128) wpc = (int)std::min<size_t>(16, (ctx->smem_optin - 64) / per_warp);   // balanced CTA queues: no wave quantisation
    if (wpc < 1) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table_log too large for shared memory");
    wpc = std::max(1, std::min(wpc, dev_opt("FSE_B200_ENC_WPC", wpc)));
    int grid = (int)std::min<size_t>((nb + wpc - 1) / wpc, (size_t)ctx->num_sms);
    if (p->n_states == 128) {
        grid = (int)std::min<size_t>(nb, (size_t)ctx->num_sms);
        // a warp codes whole blocks: keep the number of passes the CTA needs, with the fewest warps that give it
        const size_t per_cta = (nb + grid - 1) / grid, passes = (per_cta + wpc - 1) / wpc;
        wpc = (int)((per_cta + passes - 1) / passes);
        wpc = std::max(1, std::min(wpc, dev_opt("FSE_B200_ENC_WPC", wpc)));
    }
    if (p->segment_size) {
        // one table per block, its segments coded by the warps of one CTA against bank-replicated tables
        if (tlmax > SH_TL_MAX) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "segment_size needs table_log <= 11");
        const size_t fixed = sh_enc_blocks_layout<16, 16>(tlmax, 0).total, pw = ShEncStage<16>::BYTES;
        int w = (int)std::min<size_t>(std::min<size_t>(16, a.segs_per_block), (ctx->smem_optin - 64 - fixed) / pw);
        w = std::max(1, std::min(w, dev_opt("FSE_B200_SH_WARPS", w)));
        const int g = (int)std::min<size_t>(nb, (size_t)ctx->num_sms);
        Timed t(ctx, FSE_B200_K_ENCODE);
        k_encode_sh_blocks<16, 16><<<g, (w + 1) * 32, fixed + w * pw, ctx->stream>>>(a);
    } else if (global && p->n_states == 128 && tlmax <= SH_TL_MAX) {
        // one table for the job: CTA-owned, bank-replicated tables (fse_shared_enc.cuh), one CTA per SM
        // 16 rounds per chunk, 16 copies of the next-state table (32 copies / 8 rounds measured slower: DESIGN.md 4)
        const size_t fixed = sh_enc_layout<16, 16>(tlmax, 0).total, pw = ShEncStage<16>::BYTES;
        int w = (int)std::min<size_t>(16, (ctx->smem_optin - 64 - fixed) / pw);
        w = std::max(1, std::min(w, dev_opt("FSE_B200_SH_WARPS", w)));
        const int g = (int)std::min<size_t>(nb, (size_t)ctx->num_sms);
        Timed t(ctx, FSE_B200_K_ENCODE);
        k_encode_sh_global<16, 16><<<g, w * 32, fixed + w * pw, ctx->stream>>>(a);
    } else if (!global && p->n_states <= 2 && !p->flags && tlmax <= 12 && nb >= (size_t)dev_opt("FSE_B200_TPS_MIN", (int)TPS_MIN_BLOCKS) &&
               ((size_t)2 << tlmax) + 2048 <= ctx->smem_optin - 64) {
        // the reference's own one- / two-state streams, many of them: one thread per stream (fse_tps.cuh)
        const size_t wave = std::min<size_t>(nb, (size_t)dev_opt("FSE_B200_TPS_WAVE", (int)(p->n_states == 2 ? TPS_ENC_WAVE : 4 * TPS_ENC_WAVE)));
        CK(ctx->tps_enc_tab.reserve((wave << tlmax) * sizeof(uint16_t)));
        CK(ctx->tps_enc_tt.reserve(wave * 256 * sizeof(uint2)));
        CK(ctx->tps_meta.reserve(wave * sizeof(uint4)));
        Timed t(ctx, FSE_B200_K_ENCODE);
        for (size_t first = 0; first < nb; first += wave) {
            const uint32_t count = (uint32_t)std::min(wave, nb - first);
            TpsTables g{ctx->tps_enc_tab.as<uint16_t>(), ctx->tps_enc_tt.as<uint2>(), nullptr, ctx->tps_meta.as<uint4>(), (uint32_t)first, count};
            const int pg = (int)std::min<size_t>((count + wpc - 1) / wpc, (size_t)ctx->num_sms);
            k_tps_prepare_enc<<<pg, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a, g);
            const size_t set_bytes = ((size_t)2 << tlmax) + 2048;
            const uint32_t per_cta = tps_streams_per_cta(count, (size_t)ctx->num_sms, std::min<size_t>(256, (ctx->smem_optin - 64) / set_bytes));
            const uint32_t lpw = (uint32_t)std::max(1, std::min(32, dev_opt("FSE_B200_TPS_ENC_LPW", (int)((per_cta + 3) / 4))));   // four warps (c4, 37 streams: 5 / 6 / 8 / 10 / 13 lanes: 65.9 / 66.6 / 67.5 / 63.0 / 64.9 ms)
            k_tps_encode_smem<<<(count + per_cta - 1) / per_cta, ((per_cta + lpw - 1) / lpw) * 32, per_cta * set_bytes, ctx->stream>>>(a, g, per_cta, lpw);
            ctx->launches += 2;
        }
        ctx->launches--;                                     // the caller counts one
    } else {
        Timed t(ctx, FSE_B200_K_ENCODE);
        if (p->n_states == 128) k_encode128_blocks<<<grid, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a);
        else if (wide) k_encode64_blocks<<<grid, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a);
        else k_encode_blocks<<<grid, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a);
    }
    {
        Timed t(ctx, FSE_B200_K_SCAN);
        k_scan_sizes<<<1, 1024, 0, ctx->stream>>>(a.hlen, a.plen, (uint32_t)ns, reinterpret_cast<unsigned long long *>(d_offsets));
    }
    {
        Timed t(ctx, FSE_B200_K_GATHER);
        k_gather<<<(int)std::min<size_t>(ns, (size_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(
            a.scratch, stride, a.hlen, a.plen, reinterpret_cast<const unsigned long long *>(d_offsets), (uint32_t)ns, d_dst);
    }
    CK(cudaGetLastError());
    return FSE_B200_OK;
}

int fse_b200_compress_blocks(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, const fse_b200_params *p,
                             uint8_t *d_dst, size_t dst_cap, uint64_t *d_offsets, int32_t *d_status, uint64_t *h_total)
{
    int rc = fse_b200_compress_blocks_async(ctx, d_src, n, p, d_dst, dst_cap, d_offsets, d_status);
    if (rc) return rc;
    const size_t nb = num_streams(n, p);
    uint64_t total = 0;
    CK(cudaMemcpyAsync(&total, d_offsets + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (h_total) *h_total = total;
    return FSE_B200_OK;
}

int fse_b200_decompress_blocks_async(fse_b200_ctx *ctx, const uint8_t *d_comp, size_t comp_bytes, const uint64_t *d_offsets,
                                     size_t nblocks, const fse_b200_params *p, uint8_t *d_dst, size_t n, int32_t *d_status)
{
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!d_comp || !d_offsets || (!d_dst && n) || !d_status) return fail(ctx, FSE_B200_ERR_ARG, "decompress_blocks: null pointer");
    if (nblocks != num_streams(n, p)) return fail(ctx, FSE_B200_ERR_ARG, "nblocks does not match n / block_size (fse_b200_num_streams)");
    if (nblocks > 0x7fffffffull) return fail(ctx, FSE_B200_ERR_ARG, "too many blocks");
    CK(cudaSetDevice(ctx->device));
    if (nblocks == 0) return FSE_B200_OK;
    const bool global = p->table_mode == FSE_B200_TABLE_GLOBAL;
    if (global && !ctx->g_valid) return fail(ctx, FSE_B200_ERR_ARG, "global table not installed");
    const uint32_t tlmax = global ? ctx->g_log2 : tlmax_for(p);
    DecArgs a;
    a.comp = d_comp; a.comp_bytes = comp_bytes; a.offsets = reinterpret_cast<const unsigned long long *>(d_offsets);
    a.nblocks = (uint32_t)fse_b200_num_blocks(n, p->block_size); a.block_size = p->block_size; a.n = n; a.n_states = p->n_states; a.tlmax = tlmax;
    a.seg_size = p->segment_size; a.segs_per_block = p->segment_size ? p->block_size / p->segment_size : 1;
    a.dec_copies = sh_dec_copies_for(tlmax);
    a.dst = d_dst; a.status = d_status; a.global_mode = global ? 1 : 0;
    a.exhaust = 0; a.out_len = nullptr;
    a.g.log2 = ctx->g_log2; a.g.table_len = ctx->g_table_len;
    a.g.enc_table = ctx->g_enc_table.as<uint16_t>(); a.g.enc_tt = ctx->g_enc_tt.as<uint2>(); a.g.dec_table = ctx->g_dec_table.as<uint32_t>();
    if (p->n_states == 64 && tlmax > 13) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "n_states 64 needs table_log <= 13");
    if (p->segment_size) {
        // one table per block: a CTA per block, its warps decode the block's segments against a bank-replicated table
        if (tlmax > SH_TL_MAX) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "segment_size needs table_log <= 11");
        const int w = (int)std::min<uint32_t>(16, a.segs_per_block);
        // copies: as many as let two CTAs share an SM (decode needs ~32 warps in flight), at least 8
        uint32_t R = (uint32_t)dev_opt("FSE_B200_SHD_COPIES", 0);
        if (!R) {
            R = 32;
            while (R > 8 && 2 * (sh_dec_blocks_layout(tlmax, R, w).total + 1024) > ctx->smem_per_sm) R >>= 1;
        }
        a.dec_copies = R;
        const size_t smem = sh_dec_blocks_layout(tlmax, R, w).total;
        if (smem > ctx->smem_optin - 64) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table too large for shared memory");
        const int ctas = 2 * (smem + 1024) <= ctx->smem_per_sm ? 2 : 1;
        const int g = (int)std::min<size_t>(a.nblocks, (size_t)ctx->num_sms * ctas);
        Timed t(ctx, FSE_B200_K_DECODE);
        k_decode_sh_blocks<<<g, (w + 1) * 32, smem, ctx->stream>>>(a);
        CK(cudaGetLastError());
        return FSE_B200_OK;
    }
    if (global && p->n_states == 128 && tlmax <= SH_TL_MAX) {
        // CTA-owned, bank-replicated decode table (fse_shared_dec.cuh), one CTA per SM
        const int w = std::max(1, std::min(32, dev_opt("FSE_B200_SHD_WARPS", 32)));
        const int g = (int)std::min<size_t>(nblocks, (size_t)ctx->num_sms);
        Timed t(ctx, FSE_B200_K_DECODE);
        k_decode_sh_global<<<g, w * 32, sh_dec_layout(tlmax, a.dec_copies, w).total, ctx->stream>>>(a);
        CK(cudaGetLastError());
        return FSE_B200_OK;
    }
    if (p->n_states == 128) {
        if (tlmax > 13) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "n_states 128 needs table_log <= 13");
        const size_t half = ctx->smem_per_sm / 2 - 1024;
        // Two table forms.  Compact (u16 transform + u8 symbol: two look-ups per symbol, 3 * size bytes) keeps more warps
        // per SM; wide (the reference's 32-bit DecodeTransform: one look-up, 4 * size bytes) does less work per symbol.
        // Wide wins when it fits as many warps (table_log <= 10), and at table_log 11 (24 against 32 warps per SM) on inputs
        // with at least two blocks per warp slot, where the coarser whole-block passes stop mattering (c4: 8.55 against
        // 9.73 ms on 65 536 blocks, 1.232 against 1.263 ms on 8 192; c2's 4 096 blocks: 0.485 against 0.379 ms); at
        // table_log 12 (12 against 16 warps) it loses; table_log 13 has the wide form only.
        auto warps_for = [&](size_t per_warp, int &ctas_out) {
            int w = (int)std::min<size_t>(16, (std::min(half, ctx->smem_optin) - 64) / per_warp);
            ctas_out = 2;
            if (w < 1) { w = (int)std::min<size_t>(16, (ctx->smem_optin - 64) / per_warp); ctas_out = 1; }
            return w;
        };
        int ctas_c = 2, ctas_w = 2;
        const int wpc_c = tlmax <= 12 ? warps_for(dec64c_layout(tlmax).total, ctas_c) : 0;
        const int wpc_w = warps_for(dec64w_layout(tlmax).total, ctas_w);
        bool compact = wpc_c >= 1 && !(wpc_w * ctas_w >= wpc_c * ctas_c ||
                                       (tlmax <= 11 && nblocks >= (size_t)2 * ctx->num_sms * ctas_w * std::max(wpc_w, 1)));
        const int force = dev_opt("FSE_B200_DECODE128_WIDE", -1);
        if (force == 0 && wpc_c >= 1) compact = true;
        if (force == 1) compact = false;
        const size_t per_warp = compact ? dec64c_layout(tlmax).total : dec64w_layout(tlmax).total;
        // balanced CTA queues (every CTA gets nblocks / grid blocks): take every warp that fits, two CTAs per SM
        int wpc = compact ? wpc_c : wpc_w;
        int ctas = compact ? ctas_c : ctas_w;
        if (wpc < 1) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table_log too large for shared memory");
        int grid = (int)std::min<size_t>(nblocks, (size_t)ctx->num_sms * ctas);
        {   // a warp decodes whole blocks: keep the number of passes the CTA needs, with the fewest warps that give it
            // (c2: 13.8 blocks per CTA decode in 0.378 ms with 14 warps, 0.397 ms with 16)
            const size_t per_cta = (nblocks + grid - 1) / grid, passes = (per_cta + wpc - 1) / wpc;
            wpc = (int)((per_cta + passes - 1) / passes);
        }
        wpc = std::max(1, std::min(wpc, dev_opt("FSE_B200_WPC", wpc)));
        Timed t(ctx, FSE_B200_K_DECODE);
        if (compact) k_decode128c_blocks<<<grid, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a);
        else k_decode128_blocks<<<grid, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a);
        CK(cudaGetLastError());
        return FSE_B200_OK;
    }
    if (p->n_states == 64 && tlmax <= 12) {
        // compact tables, two CTAs per SM, each with half of the SM's shared memory (wide entries measured 0.70 vs 0.49 ms on c2)
        const size_t half = ctx->smem_per_sm / 2 - 1024;            // 1 KiB per CTA is reserved by the runtime
        const size_t per_warp = dec64c_layout(tlmax).total;
        int wpc = pick_warps(nblocks, ctx->num_sms * 2, per_warp, std::min(half, ctx->smem_optin), 16);
        int ctas = 2;
        if (wpc < 1) { wpc = pick_warps(nblocks, ctx->num_sms, per_warp, ctx->smem_optin, 16); ctas = 1; }
        if (wpc < 1) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table_log too large for shared memory");
        int grid = (int)std::min<size_t>((nblocks + wpc - 1) / wpc, (size_t)ctx->num_sms * ctas);
        Timed t(ctx, FSE_B200_K_DECODE);
        k_decode64c_blocks<<<grid, wpc * 32, (size_t)wpc * per_warp, ctx->stream>>>(a);
    } else {
        const DecLayout lay = dec_layout(tlmax);
        int wpc = pick_warps(nblocks, ctx->num_sms, lay.total, ctx->smem_optin, 16);
        if (wpc < 1) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table_log too large for shared memory");
        int grid = (int)std::min<size_t>((nblocks + wpc - 1) / wpc, (size_t)ctx->num_sms);
        if (!global && !a.exhaust && p->n_states <= 2 && tlmax <= 12 &&
            nblocks >= (size_t)dev_opt("FSE_B200_TPS_MIN", (int)TPS_MIN_BLOCKS) && ((size_t)4 << tlmax) <= ctx->smem_optin - 64) {
            // the reference's own one- / two-state streams, many of them: one thread per stream (fse_tps.cuh)
            CK(ctx->tps_dec_tab.reserve((nblocks << tlmax) * sizeof(uint32_t)));
            CK(ctx->tps_meta.reserve(nblocks * sizeof(uint4)));
            TpsTables g{nullptr, nullptr, ctx->tps_dec_tab.as<uint32_t>(), ctx->tps_meta.as<uint4>(), 0u, (uint32_t)nblocks};
            Timed t(ctx, FSE_B200_K_DECODE);
            k_tps_prepare_dec<<<grid, wpc * 32, (size_t)wpc * lay.total, ctx->stream>>>(a, g);
            // 16-bit entries + symbol bytes (3 bytes per cell, table_log <= 11) when the smaller sets save a round of CTAs
            // (c4: 37 streams per SM instead of 28, 12 rounds instead of 16: 83 -> 72 ms); the 32-bit DecodeTransform
            // otherwise (fewer instructions per symbol).  Lanes per warp: about eight warps per CTA (measured 2 .. 8).
            const size_t fit_w = std::min<size_t>(256, (ctx->smem_optin - 64) / ((size_t)4 << tlmax));
            const size_t fit_c = std::min<size_t>(256, (ctx->smem_optin - 64) / ((size_t)3 << tlmax));
            auto rounds = [&](size_t fit) { return fit ? (nblocks + ctx->num_sms * fit - 1) / (ctx->num_sms * fit) : (size_t)-1; };
            const bool compact = dev_opt("FSE_B200_TPS_COMPACT", tlmax <= 11 && rounds(fit_c) < rounds(fit_w));
            const size_t set_bytes = (size_t)(compact ? 3 : 4) << tlmax;
            const uint32_t per_cta = tps_streams_per_cta(nblocks, (size_t)ctx->num_sms, compact ? fit_c : fit_w);
            const uint32_t lpw = (uint32_t)std::max(1, std::min(32, dev_opt("FSE_B200_TPS_LPW", (int)((per_cta + 7) / 8))));
            const unsigned dgrid = (unsigned)((nblocks + per_cta - 1) / std::max(per_cta, 1u)), dthreads = ((per_cta + lpw - 1) / lpw) * 32;
            if (compact) k_tps_decode_smem<true><<<dgrid, dthreads, per_cta * set_bytes, ctx->stream>>>(a, g, per_cta, lpw);
            else k_tps_decode_smem<false><<<dgrid, dthreads, per_cta * set_bytes, ctx->stream>>>(a, g, per_cta, lpw);
            ctx->launches++;
            CK(cudaGetLastError());
            return FSE_B200_OK;
        }
        Timed t(ctx, FSE_B200_K_DECODE);
        if (p->n_states == 64) k_decode64_blocks<<<grid, wpc * 32, (size_t)wpc * lay.total, ctx->stream>>>(a);
        else k_decode_blocks<<<grid, wpc * 32, (size_t)wpc * lay.total, ctx->stream>>>(a);
    }
    CK(cudaGetLastError());
    return FSE_B200_OK;
}

int fse_b200_decompress_blocks(fse_b200_ctx *ctx, const uint8_t *d_comp, size_t comp_bytes, const uint64_t *d_offsets,
                               size_t nblocks, const fse_b200_params *p, uint8_t *d_dst, size_t n, int32_t *d_status)
{
    int rc = fse_b200_decompress_blocks_async(ctx, d_comp, comp_bytes, d_offsets, nblocks, p, d_dst, n, d_status);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

int fse_b200_decompress_exhaust(fse_b200_ctx *ctx, const uint8_t *d_comp, size_t comp_bytes, const uint64_t *d_offsets,
                                size_t nblocks, const fse_b200_params *p, uint8_t *d_dst, uint32_t *d_out_len, int32_t *d_status)
{
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!d_comp || !d_offsets || !d_dst || !d_out_len || !d_status) return fail(ctx, FSE_B200_ERR_ARG, "decompress_exhaust: null pointer");
    if (p->table_mode != FSE_B200_TABLE_PER_BLOCK || p->segment_size) return fail(ctx, FSE_B200_ERR_ARG, "decompress_exhaust: per-block tables, no segments");
    if (p->n_states > 32) return fail(ctx, FSE_B200_ERR_ARG, "decompress_exhaust: n_states <= 32");
    if (nblocks > 0x7fffffffull) return fail(ctx, FSE_B200_ERR_ARG, "too many blocks");
    CK(cudaSetDevice(ctx->device));
    if (nblocks == 0) return FSE_B200_OK;
    const uint32_t tlmax = p->table_log ? p->table_log : 11u;
    DecArgs a;
    a.comp = d_comp; a.comp_bytes = comp_bytes; a.offsets = reinterpret_cast<const unsigned long long *>(d_offsets);
    a.nblocks = (uint32_t)nblocks; a.block_size = p->block_size; a.n = nblocks * (size_t)p->block_size;
    a.n_states = p->n_states; a.tlmax = tlmax;
    a.dst = d_dst; a.status = d_status; a.global_mode = 0; a.exhaust = 1; a.out_len = d_out_len;
    a.seg_size = 0; a.segs_per_block = 1; a.dec_copies = 0;
    a.g.log2 = 0; a.g.table_len = 0; a.g.enc_table = nullptr; a.g.enc_tt = nullptr; a.g.dec_table = nullptr;
    const DecLayout lay = dec_layout(tlmax);
    int wpc = pick_warps(nblocks, ctx->num_sms, lay.total, ctx->smem_optin, 16);
    if (wpc < 1) return fail(ctx, FSE_B200_ERR_UNSUPPORTED, "table_log too large for shared memory");
    int grid = (int)std::min<size_t>((nblocks + wpc - 1) / wpc, (size_t)ctx->num_sms);
    k_decode_blocks<<<grid, wpc * 32, (size_t)wpc * lay.total, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

// ---------------------------------------------------------------------------------- host buffers
// Large host buffers are processed in chunks of whole blocks so that the H2D copy of chunk i+1, the kernels of
// chunk i and the D2H copy of chunk i-1 overlap (PCIe is full duplex): three streams, events between them.  Every
// small read-back (per-chunk totals, the index, the status words) lands in the context's pinned buffer, so that no
// copy inside the loop is a staged synchronous one.  "Streams" below are the entries of the index: blocks, or
// segments when p->segment_size > 0.

static int worst_status(const int32_t *st, size_t nb)
{
    for (size_t i = 0; i < nb; i++) if (st[i] < 0) return FSE_B200_ERR_BLOCK;
    return FSE_B200_OK;
}

static const size_t PIPE_CHUNK_BYTES = 32u << 20;

// blocks a chunk must hold: one- and two-state streams are coded one thread per stream (fse_tps.cuh), whose throughput
// grows with the streams per call until a round of CTAs is full, so their chunks are that large
static size_t pipe_min_blocks(const fse_b200_params *p)
{
    return (p->n_states <= 2 && p->table_mode == FSE_B200_TABLE_PER_BLOCK && !p->flags) ? (size_t)TPS_PIPE_BLOCKS : 1;
}

static int pipe_setup(fse_b200_ctx *ctx, size_t nchunks)
{
    if (!ctx->s_in) CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    while (ctx->pipe_ev.size() < 2 * nchunks + 2) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    return FSE_B200_OK;
}

int fse_b200_compress_host(fse_b200_ctx *ctx, const uint8_t *h_src, size_t n, const fse_b200_params *p, uint8_t *h_dst,
                           size_t dst_cap, uint64_t *h_offsets, int32_t *h_status, uint64_t *h_total)
{
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if ((!h_src && n) || !h_dst || !h_total) return fail(ctx, FSE_B200_ERR_ARG, "compress_host: null pointer");
    CK(cudaSetDevice(ctx->device));
    const size_t bs = p->block_size, S = bs / stream_bytes(p);
    const size_t ns = num_streams(n, p);
    // chunks of whole blocks; small inputs are one chunk
    const size_t cblocks = (n > 2 * PIPE_CHUNK_BYTES && bs <= PIPE_CHUNK_BYTES) ? std::max<size_t>(pipe_min_blocks(p), PIPE_CHUNK_BYTES / bs)
                                                                                 : std::max<size_t>(1, fse_b200_num_blocks(n, p->block_size));
    const size_t cbytes = cblocks * bs, cstreams = cblocks * S;
    const size_t nchunks = std::max<size_t>(1, (n + cbytes - 1) / cbytes);
    const size_t cbound = fse_b200_compress_blocks_bound(std::min(cbytes, n), p);
    rc = pipe_setup(ctx, nchunks);
    if (rc) return rc;
    CK(ctx->stage_in.reserve(n + 16));
    CK(ctx->stage_out.reserve(nchunks * cbound));
    CK(ctx->stage_off.reserve(nchunks * (cstreams + 1) * 8));
    CK(ctx->stage_status.reserve((ns + 1) * 4));
    // pinned: totals[nchunks] | loc[nchunks][cstreams + 1] | status[ns]
    const size_t pin_loc = nchunks * 8, pin_st = pin_loc + nchunks * (cstreams + 1) * 8;
    CK(ctx->pin.reserve(pin_st + (ns + 1) * 4));
    uint64_t *h_tot = reinterpret_cast<uint64_t *>(ctx->pin.p);
    uint64_t *loc = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(ctx->pin.p) + pin_loc);
    int32_t *st = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(ctx->pin.p) + pin_st);
    // an event orders the copy streams after whatever the caller queued on the compute stream
    CK(cudaEventRecord(ctx->pipe_ev[2 * nchunks], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_in, ctx->pipe_ev[2 * nchunks], 0));
    for (size_t c = 0; c < nchunks && n; c++) {
        size_t o = c * cbytes, len = std::min(cbytes, n - o);
        CK(cudaMemcpyAsync(ctx->stage_in.as<uint8_t>() + o, h_src + o, len, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ctx->pipe_ev[2 * c], ctx->s_in));
    }
    for (size_t c = 0; c < nchunks && n; c++) {
        size_t o = c * cbytes, len = std::min(cbytes, n - o), cns = num_streams(len, p);
        uint64_t *d_off = ctx->stage_off.as<uint64_t>() + c * (cstreams + 1);
        CK(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 * c], 0));
        rc = fse_b200_compress_blocks_async(ctx, ctx->stage_in.as<uint8_t>() + o, len, p, ctx->stage_out.as<uint8_t>() + c * cbound,
                                            cbound, d_off, ctx->stage_status.as<int32_t>() + c * cstreams);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h_tot + c, d_off + cns, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->pipe_ev[2 * c + 1], ctx->stream));
    }
    uint64_t run = 0;
    int ret = FSE_B200_OK;
    for (size_t c = 0; c < nchunks && n; c++) {
        size_t o = c * cbytes, len = std::min(cbytes, n - o), cns = num_streams(len, p);
        CK(cudaEventSynchronize(ctx->pipe_ev[2 * c + 1]));
        const uint64_t tot = h_tot[c];
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->pipe_ev[2 * c + 1], 0));
        if (run + tot > dst_cap) ret = FSE_B200_ERR_CAPACITY;
        else if (tot) CK(cudaMemcpyAsync(h_dst + run, ctx->stage_out.as<uint8_t>() + c * cbound, tot, cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaMemcpyAsync(loc + c * (cstreams + 1), ctx->stage_off.as<uint64_t>() + c * (cstreams + 1), (cns + 1) * 8,
                           cudaMemcpyDeviceToHost, ctx->s_out));
        run += tot;
    }
    if (ns) CK(cudaMemcpyAsync(st, ctx->stage_status.p, ns * 4, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(ctx->stream));
    *h_total = run;
    if (ret) return fail(ctx, ret, "compress_host: dst_cap too small");
    if (h_offsets) {
        uint64_t base = 0;
        for (size_t c = 0; c < nchunks && n; c++) {
            size_t o = c * cbytes, len = std::min(cbytes, n - o), cns = num_streams(len, p);
            const uint64_t *l = loc + c * (cstreams + 1);
            for (size_t b = 0; b < cns; b++) h_offsets[c * cstreams + b] = base + l[b];
            base += l[cns];
        }
        h_offsets[ns] = base;
    }
    if (h_status && ns) memcpy(h_status, st, ns * 4);
    return worst_status(st, ns);
}

int fse_b200_decompress_host(fse_b200_ctx *ctx, const uint8_t *h_comp, size_t comp_bytes, const uint64_t *h_offsets,
                             size_t nblocks, const fse_b200_params *p, uint8_t *h_dst, size_t n, int32_t *h_status)
{
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if (!h_comp || !h_offsets || (!h_dst && n)) return fail(ctx, FSE_B200_ERR_ARG, "decompress_host: null pointer");
    CK(cudaSetDevice(ctx->device));
    const size_t ns = num_streams(n, p);
    if (nblocks != ns) return fail(ctx, FSE_B200_ERR_ARG, "nblocks does not match n / block_size (fse_b200_num_streams)");
    if (ns > 0x7fffffffull) return fail(ctx, FSE_B200_ERR_ARG, "too many blocks");
    if (ns == 0) return FSE_B200_OK;
    for (size_t b = 0; b < ns; b++)
        if (h_offsets[b + 1] < h_offsets[b]) return fail(ctx, FSE_B200_ERR_ARG, "decompress_host: offsets not monotone");
    if (h_offsets[ns] > comp_bytes) return fail(ctx, FSE_B200_ERR_ARG, "decompress_host: offsets exceed comp_bytes");
    const size_t bs = p->block_size, S = bs / stream_bytes(p);
    const size_t nb = fse_b200_num_blocks(n, p->block_size);
    const size_t cblocks = (n > 2 * PIPE_CHUNK_BYTES && bs <= PIPE_CHUNK_BYTES) ? std::max<size_t>(pipe_min_blocks(p), PIPE_CHUNK_BYTES / bs) : nb;
    const size_t cbytes = cblocks * bs, cstreams = cblocks * S;
    const size_t nchunks = (nb + cblocks - 1) / cblocks;
    rc = pipe_setup(ctx, nchunks);
    if (rc) return rc;
    CK(ctx->stage_out.reserve(comp_bytes + 16));
    CK(ctx->stage_in.reserve(n + 16));
    CK(ctx->stage_off.reserve((ns + 1) * 8));
    CK(ctx->stage_status.reserve((ns + 1) * 4));
    CK(ctx->pin.reserve((ns + 1) * 4));
    int32_t *st = reinterpret_cast<int32_t *>(ctx->pin.p);
    CK(cudaEventRecord(ctx->pipe_ev[2 * nchunks], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->s_in, ctx->pipe_ev[2 * nchunks], 0));
    CK(cudaMemcpyAsync(ctx->stage_off.p, h_offsets, (ns + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
    for (size_t c = 0; c < nchunks; c++) {
        const size_t s0 = c * cstreams, s1 = std::min(ns, s0 + cstreams);
        const uint64_t o0 = h_offsets[s0], o1 = h_offsets[s1];
        if (o1 > o0) CK(cudaMemcpyAsync(ctx->stage_out.as<uint8_t>() + o0, h_comp + o0, o1 - o0, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ctx->pipe_ev[2 * c], ctx->s_in));
    }
    for (size_t c = 0; c < nchunks; c++) {
        const size_t s0 = c * cstreams, s1 = std::min(ns, s0 + cstreams);
        const size_t o = c * cbytes, len = std::min(cbytes, n - o);
        CK(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 * c], 0));
        rc = fse_b200_decompress_blocks_async(ctx, ctx->stage_out.as<uint8_t>(), comp_bytes, ctx->stage_off.as<uint64_t>() + s0, s1 - s0,
                                              p, ctx->stage_in.as<uint8_t>() + o, len, ctx->stage_status.as<int32_t>() + s0);
        if (rc) return rc;
        CK(cudaEventRecord(ctx->pipe_ev[2 * c + 1], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->pipe_ev[2 * c + 1], 0));
        CK(cudaMemcpyAsync(h_dst + o, ctx->stage_in.as<uint8_t>() + o, len, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CK(cudaMemcpyAsync(st, ctx->stage_status.p, ns * 4, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(ctx->stream));
    if (h_status) memcpy(h_status, st, ns * 4);
    return worst_status(st, ns);
}

// ---------------------------------------------------------------------------------- frame (container)

namespace {
struct FrameHeader {            // 56 bytes, little endian
    uint32_t magic;
    uint16_t version, n_states;
    uint32_t block_size, table_log, table_mode, global_header_bytes;
    uint64_t n, nstreams, payload_bytes;
    uint32_t segment_size, flags;
};
static_assert(sizeof(FrameHeader) == 56, "frame header layout");
size_t pad8(size_t v) { return (v + 7) & ~(size_t)7; }
}  // namespace

size_t fse_b200_frame_bound(size_t n, const fse_b200_params *p)
{
    if (!p || !p->block_size) return 0;
    return sizeof(FrameHeader) + 512 + (num_streams(n, p) + 1) * 8 + fse_b200_compress_blocks_bound(n, p);
}

// Everything in a frame header is untrusted: each term is checked against what is left of the frame before it is
// used, so that no sum can wrap, and the parameters must be ones the library could have written.
int fse_b200_frame_info(const uint8_t *h_frame, size_t frame_bytes, fse_b200_params *p_out, size_t *n_out)
{
    if (!h_frame || frame_bytes < sizeof(FrameHeader)) return FSE_B200_ERR_ARG;
    FrameHeader h;
    memcpy(&h, h_frame, sizeof(h));
    if (h.magic != FSE_B200_FRAME_MAGIC || h.version != 2) return FSE_B200_ERR_ARG;
    fse_b200_params p;
    p.block_size = h.block_size; p.table_log = h.table_log; p.n_states = h.n_states; p.table_mode = h.table_mode;
    p.segment_size = h.segment_size; p.flags = h.flags;
    fse_b200_ctx scratch_ctx;                               // only its error string is touched
    if (check_params(&scratch_ctx, &p)) return FSE_B200_ERR_ARG;
    if (h.global_header_bytes > 512 || (p.table_mode == FSE_B200_TABLE_GLOBAL) != (h.global_header_bytes != 0)) return FSE_B200_ERR_ARG;
    if (h.n > ((uint64_t)1 << 62)) return FSE_B200_ERR_ARG;
    if (h.nstreams != num_streams((size_t)h.n, &p) || h.nstreams > 0x7fffffffull) return FSE_B200_ERR_ARG;
    size_t left = frame_bytes - sizeof(FrameHeader);
    const size_t gh = pad8(h.global_header_bytes);
    if (gh > left) return FSE_B200_ERR_CAPACITY;
    left -= gh;
    const size_t idx = ((size_t)h.nstreams + 1) * 8;        // nstreams < 2^31: cannot wrap
    if (idx > left) return FSE_B200_ERR_CAPACITY;
    left -= idx;
    if (h.payload_bytes > left) return FSE_B200_ERR_CAPACITY;
    if (p_out) *p_out = p;
    if (n_out) *n_out = (size_t)h.n;
    return FSE_B200_OK;
}

int fse_b200_frame_compress_host(fse_b200_ctx *ctx, const uint8_t *h_src, size_t n, const fse_b200_params *p,
                                 uint8_t *h_frame, size_t frame_cap, size_t *h_frame_bytes)
{
    int rc = check_params(ctx, p);
    if (rc) return rc;
    if ((!h_src && n) || !h_frame || !h_frame_bytes) return fail(ctx, FSE_B200_ERR_ARG, "frame_compress_host: null pointer");
    CK(cudaSetDevice(ctx->device));
    const size_t ns = num_streams(n, p);
    FrameHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = FSE_B200_FRAME_MAGIC; h.version = 2; h.n_states = (uint16_t)p->n_states;
    h.block_size = p->block_size; h.table_log = p->table_log; h.table_mode = p->table_mode;
    h.segment_size = p->segment_size; h.flags = p->flags;
    h.n = n; h.nstreams = ns;
    uint8_t ghdr[512];
    size_t gbytes = 0;
    if (p->table_mode == FSE_B200_TABLE_GLOBAL) {
        // one table for the whole frame: histogram of the input on the device, normalise, keep the header
        CK(ctx->stage_in.reserve(n + 16));
        CK(ctx->g_meta.reserve(64));
        CK(ctx->misc.reserve(256 * 8 + 16));
        CK(cudaMemcpyAsync(ctx->stage_in.p, h_src, n, cudaMemcpyHostToDevice, ctx->stream));
        rc = hist_global_async(ctx, ctx->stage_in.as<uint8_t>(), n, ctx->misc.as<uint64_t>());
        if (rc) return rc;
        gbytes = sizeof(ghdr);
        uint32_t l2 = 0;
        rc = fse_b200_set_global_table(ctx, ctx->misc.as<uint64_t>(), p->table_log, ghdr, &gbytes, &l2);
        if (rc) return rc;
        h.global_header_bytes = (uint32_t)gbytes;
    }
    const size_t off_pos = sizeof(FrameHeader) + pad8(gbytes);
    const size_t pay_pos = off_pos + (ns + 1) * 8;
    if (frame_cap < pay_pos) return fail(ctx, FSE_B200_ERR_CAPACITY, "frame_compress_host: frame_cap too small");
    std::vector<uint64_t> off(ns + 1);
    uint64_t total = 0;
    rc = fse_b200_compress_host(ctx, h_src, n, p, h_frame + pay_pos, frame_cap - pay_pos, off.data(), nullptr, &total);
    if (rc != FSE_B200_OK && rc != FSE_B200_ERR_BLOCK) return rc;
    h.payload_bytes = total;
    memcpy(h_frame, &h, sizeof(h));
    if (gbytes) { memset(h_frame + sizeof(FrameHeader), 0, pad8(gbytes)); memcpy(h_frame + sizeof(FrameHeader), ghdr, gbytes); }
    memcpy(h_frame + off_pos, off.data(), (ns + 1) * 8);
    *h_frame_bytes = pay_pos + total;
    return rc;                                              // FSE_B200_ERR_BLOCK: a block could not be coded, the frame cannot be decoded
}

int fse_b200_frame_decompress_host(fse_b200_ctx *ctx, const uint8_t *h_frame, size_t frame_bytes, uint8_t *h_dst,
                                   size_t dst_cap, size_t *h_n)
{
    if (!ctx || !h_n) return FSE_B200_ERR_ARG;
    fse_b200_params p;
    size_t n = 0;
    int rc = fse_b200_frame_info(h_frame, frame_bytes, &p, &n);
    if (rc) return fail(ctx, rc, "frame_decompress_host: bad frame");
    rc = check_params(ctx, &p);
    if (rc) return rc;
    if (n > dst_cap || (!h_dst && n)) return fail(ctx, FSE_B200_ERR_CAPACITY, "frame_decompress_host: dst_cap too small");
    FrameHeader h;
    memcpy(&h, h_frame, sizeof(h));
    if (p.table_mode == FSE_B200_TABLE_GLOBAL) {
        uint32_t l2 = 0;
        rc = fse_b200_set_global_table_from_header(ctx, h_frame + sizeof(FrameHeader), h.global_header_bytes, &l2);
        if (rc) return rc;
    }
    const size_t off_pos = sizeof(FrameHeader) + pad8(h.global_header_bytes);
    std::vector<uint64_t> off((size_t)h.nstreams + 1);      // bounded by the frame size (fse_b200_frame_info)
    memcpy(off.data(), h_frame + off_pos, ((size_t)h.nstreams + 1) * 8);     // the frame may be unaligned
    if (off[h.nstreams] != h.payload_bytes) return fail(ctx, FSE_B200_ERR_ARG, "frame_decompress_host: offsets do not match payload size");
    *h_n = n;
    return fse_b200_decompress_host(ctx, h_frame + off_pos + ((size_t)h.nstreams + 1) * 8, (size_t)h.payload_bytes, off.data(),
                                    (size_t)h.nstreams, &p, h_dst, n, nullptr);
}

// ---------------------------------------------------------------------------------- generators

static uint32_t build_lut(int kind, std::vector<uint8_t> &lut)
{
    static const uint8_t TEXT_RANKS[96] = {
        0x20,0x65,0x74,0x61,0x6f,0x69,0x6e,0x73,0x68,0x72,0x64,0x6c,0x63,0x75,0x6d,0x77,
        0x66,0x67,0x79,0x70,0x62,0x76,0x6b,0x6a,0x78,0x71,0x7a,0x45,0x54,0x41,0x4f,0x49,
        0x4e,0x53,0x48,0x52,0x44,0x4c,0x43,0x55,0x4d,0x57,0x46,0x47,0x59,0x50,0x42,0x56,
        0x4b,0x4a,0x58,0x51,0x5a,0x30,0x31,0x32,0x33,0x34,0x35,0x36,0x37,0x38,0x39,0x2e,
        0x2c,0x3b,0x3a,0x27,0x22,0x21,0x3f,0x2d,0x28,0x29,0x0a,0x09,0x2f,0x26,0x25,0x24,
        0x23,0x40,0x2a,0x2b,0x3c,0x3d,0x3e,0x5b,0x5d,0x5f,0x7b,0x7d,0x7c,0x7e,0x5e,0x60};
    lut.clear();
    if (kind == 0) {        // the crate's gen_sequence(0.2) LUT, src/lib.rs:255-270
        size_t remaining = 4096;
        uint8_t s = 0;
        while (remaining > 0) {
            size_t k = (size_t)((double)remaining * 0.2);
            if (k < 1) k = 1;
            lut.insert(lut.end(), k, s);
            s++;
            remaining -= k;
        }
    } else if (kind == 1) { // Zipf(1) over 96 printable bytes
        uint64_t w[96], W = 0, cnt[96], used = 0;
        for (int r = 0; r < 96; r++) { w[r] = (1ull << 20) / (uint64_t)(r + 1); W += w[r]; }
        for (int r = 0; r < 96; r++) { cnt[r] = (65536ull * w[r]) / W; used += cnt[r]; }
        cnt[0] += 65536 - used;
        for (int r = 0; r < 96; r++) lut.insert(lut.end(), (size_t)cnt[r], TEXT_RANKS[r]);
    } else if (kind == 2) { // four symbols, ~.90/.05/.03/.02
        const unsigned c[4] = {3686, 205, 123, 82};
        for (int s = 0; s < 4; s++) lut.insert(lut.end(), c[s], (uint8_t)s);
    } else {
        for (int i = 0; i < 256; i++) lut.push_back((uint8_t)i);
    }
    return (uint32_t)lut.size();
}

int fse_b200_generate(fse_b200_ctx *ctx, int kind, uint64_t seed, uint64_t first_index, uint8_t *d_dst, size_t n)
{
    if (!ctx || kind < 0 || kind > 3 || (!d_dst && n)) return fail(ctx, FSE_B200_ERR_ARG, "generate: bad argument");
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return FSE_B200_OK;
    if (!ctx->lut_len[kind]) {
        std::vector<uint8_t> lut;
        uint32_t len = build_lut(kind, lut);
        CK(ctx->lut[kind].reserve(len));
        CK(cudaMemcpy(ctx->lut[kind].p, lut.data(), len, cudaMemcpyHostToDevice));
        ctx->lut_len[kind] = len;
    }
    k_generate<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->lut[kind].as<uint8_t>(), ctx->lut_len[kind] - 1, seed, first_index, d_dst, n);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return FSE_B200_OK;
}

}  // extern "C"
