// fse_kernels.cuh -- the sm_100a kernels of the FSE path.  One warp per block / per table.
#pragma once
#include "fse_device.cuh"

namespace fsed {

// ------------------------------------------------------------------------------------------
// Per-warp shared-memory slices
// ------------------------------------------------------------------------------------------
struct EncLayout {
    uint32_t tab;    // uint16[size]      next-state table (also the posmap scratch of warp_spread)
    uint32_t tt;     // uint2[256]        symbol transforms
    uint32_t work;   // build: counts u32[256] | norm i32[256] | cum u32[256] | spread u8[size]
                     // encode: fld u32[32*32]
    uint32_t rows;   // uint32[32*17]     lane bit strings
    uint32_t total;
};
__host__ __device__ inline EncLayout enc_layout(uint32_t tlmax)
{
    EncLayout l;
    uint32_t size = 1u << tlmax;
    l.tab = 0;
    l.tt = l.tab + size * 2;
    l.work = l.tt + 2048;
    uint32_t build = 3072 + size, enc = 4096;
    l.rows = l.work + (build > enc ? build : enc);
    l.total = (l.rows + ROWS_WORDS * 4 + 15u) & ~15u;
    return l;
}

struct DecLayout {
    uint32_t tab;    // uint32[size]      decode entries (also the posmap scratch)
    uint32_t norm;   // int32[256]
    uint32_t ctr;    // uint32[256]       cum (unused output of warp_spread) then symbol_next
    uint32_t spread; // uint8[size]
    uint32_t ring;   // uint32[256]       payload staging ring
    uint32_t total;
};
__host__ __device__ inline DecLayout dec_layout(uint32_t tlmax)
{
    DecLayout l;
    uint32_t size = 1u << tlmax;
    l.tab = 0;
    l.norm = size * 4;
    l.ctr = l.norm + 1024;
    l.spread = l.ctr + 1024;
    l.ring = (l.spread + size + 15u) & ~15u;
    l.total = l.ring + 1024;
    return l;
}

// A table shared by every block (FSE_B200_TABLE_GLOBAL)
struct GlobalTable {
    uint32_t log2;
    uint32_t table_len;
    const uint16_t *enc_table;  // [1 << log2]
    const uint2 *enc_tt;        // [256]
    const uint32_t *dec_table;  // [1 << log2]
};

// ------------------------------------------------------------------------------------------
// K1: Histogram::new per block (src/histogram.rs:18-66).  One CTA of 2 warps per block; every
// lane owns a private column of 256 32-bit counters (bank == lane: no conflicts, no atomics);
// four bytes are counted per step with the duplicate increments resolved in registers.
// ------------------------------------------------------------------------------------------
constexpr int HIST_WARPS = 2;
constexpr int HIST_SMEM = HIST_WARPS * 256 * 32 * 4;

__device__ __forceinline__ void hist_word(uint32_t *cnt, uint32_t w)
{
    uint32_t b0 = w & 0xff, b1 = (w >> 8) & 0xff, b2 = (w >> 16) & 0xff, b3 = w >> 24;
    uint32_t c0 = cnt[b0 << 5], c1 = cnt[b1 << 5], c2 = cnt[b2 << 5], c3 = cnt[b3 << 5];
    uint32_t i1 = (b1 == b0), i2 = (b2 == b0) + (b2 == b1), i3 = (b3 == b0) + (b3 == b1) + (b3 == b2);
    cnt[b0 << 5] = c0 + 1;  // later stores win: program order is ascending multiplicity
    cnt[b1 << 5] = c1 + 1 + i1;
    cnt[b2 << 5] = c2 + 1 + i2;
    cnt[b3 << 5] = c3 + 1 + i3;
}

__global__ void __launch_bounds__(HIST_WARPS * 32)
k_hist_blocks(const uint8_t *__restrict__ src, size_t n, uint32_t block_size, uint32_t nblocks,
              uint32_t *__restrict__ counts, uint32_t *__restrict__ table_len)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *cnt_all = reinterpret_cast<uint32_t *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *cnt = cnt_all + warp * 8192 + lane;  // counter(bin) = cnt[bin << 5]
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        for (int i = 0; i < 256; i++) cnt[i << 5] = 0;
        const size_t off = (size_t)b * block_size;
        const uint32_t bn = (uint32_t)min((size_t)block_size, n - off);
        const uint8_t *p = src + off;
        // bytes before the first 16-byte boundary and after the last one go one at a time
        uint32_t head = (uint32_t)((16 - ((uintptr_t)p & 15)) & 15);
        if (head > bn) head = bn;
        uint32_t nvec = (bn - head) >> 4;
        uint32_t tail0 = head + (nvec << 4);
        if (tid < (int)head) cnt[(uint32_t)p[tid] << 5] += 1;
        if (tail0 + tid < bn) cnt[(uint32_t)p[tail0 + tid] << 5] += 1;  // tail < 16 <= 64 threads
        const uint4 *v = reinterpret_cast<const uint4 *>(p + head);
        uint32_t i = tid;
        for (; i + 3 * HIST_WARPS * 32 < nvec; i += 4 * HIST_WARPS * 32) {  // 4 loads in flight
            uint4 x0 = __ldg(v + i), x1 = __ldg(v + i + HIST_WARPS * 32);
            uint4 x2 = __ldg(v + i + 2 * HIST_WARPS * 32), x3 = __ldg(v + i + 3 * HIST_WARPS * 32);
            hist_word(cnt, x0.x); hist_word(cnt, x0.y); hist_word(cnt, x0.z); hist_word(cnt, x0.w);
            hist_word(cnt, x1.x); hist_word(cnt, x1.y); hist_word(cnt, x1.z); hist_word(cnt, x1.w);
            hist_word(cnt, x2.x); hist_word(cnt, x2.y); hist_word(cnt, x2.z); hist_word(cnt, x2.w);
            hist_word(cnt, x3.x); hist_word(cnt, x3.y); hist_word(cnt, x3.z); hist_word(cnt, x3.w);
        }
        for (; i < nvec; i += HIST_WARPS * 32) {
            uint4 x = __ldg(v + i);
            hist_word(cnt, x.x); hist_word(cnt, x.y); hist_word(cnt, x.z); hist_word(cnt, x.w);
        }
        __syncthreads();
        // merge: thread t sums bins 4t..4t+3 over all 64 private columns (rotated => bank == lane)
        uint32_t s[4] = {0, 0, 0, 0};
        for (int w = 0; w < HIST_WARPS; w++)
            for (int l = 0; l < 32; l++) {
                int col = (l + tid) & 31;
#pragma unroll
                for (int q = 0; q < 4; q++) s[q] += cnt_all[w * 8192 + ((tid * 4 + q) << 5) + col];
            }
        *reinterpret_cast<uint4 *>(counts + (size_t)b * 256 + tid * 4) = make_uint4(s[0], s[1], s[2], s[3]);
        if (table_len) {
            int hi = -1;
#pragma unroll
            for (int q = 0; q < 4; q++) if (s[q]) hi = tid * 4 + q;
#pragma unroll
            for (int d = 16; d; d >>= 1) hi = max(hi, __shfl_xor_sync(FULL, hi, d));
            __shared__ int hiw[HIST_WARPS];
            if (lane == 0) hiw[warp] = hi;
            __syncthreads();
            if (tid == 0) {
                int h = hiw[0];
                for (int w = 1; w < HIST_WARPS; w++) h = max(h, hiw[w]);
                table_len[b] = (uint32_t)(h < 0 ? 0 : h) + 1;
            }
        }
        __syncthreads();
    }
}

// sum of per-block histograms into uint64[256] (global-table mode)
__global__ void k_hist_reduce(const uint32_t *__restrict__ counts, uint32_t nblocks, unsigned long long *out)
{
    uint32_t bin = threadIdx.x;  // 256 threads
    unsigned long long s = 0;
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) s += counts[(size_t)b * 256 + bin];
    if (s) atomicAdd(out + bin, s);
}

// sum of the histograms of the `k` pieces of each block (large blocks are counted piecewise so that few blocks still fill
// the machine: launch_hist) + table_len (src/histogram.rs:52-59).  One CTA of 256 threads per block.
__global__ void __launch_bounds__(256)
k_hist_sum_pieces(const uint32_t *__restrict__ piece_counts, uint32_t npieces, uint32_t k, uint32_t nblocks,
                  uint32_t *__restrict__ counts, uint32_t *__restrict__ table_len)
{
    __shared__ int s_hi[8];
    const uint32_t bin = threadIdx.x;
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const uint32_t p0 = b * k, p1 = min(npieces, p0 + k);
        uint32_t s = 0;
        for (uint32_t p = p0; p < p1; p++) s += piece_counts[(size_t)p * 256 + bin];
        counts[(size_t)b * 256 + bin] = s;
        if (table_len) {
            int hi = s ? (int)bin : -1;
#pragma unroll
            for (int d = 16; d; d >>= 1) hi = max(hi, __shfl_xor_sync(FULL, hi, d));
            if ((bin & 31) == 0) s_hi[bin >> 5] = hi;
            __syncthreads();
            if (bin == 0) {
                int m = -1;
                for (int i = 0; i < 8; i++) m = max(m, s_hi[i]);
                table_len[b] = (uint32_t)(m < 0 ? 0 : m) + 1;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2-K4 fused: per block  normalise -> header -> encode table -> reverse-order N-state encode.
// One warp per block; CTAs are just containers of independent warps.
// Block scratch (stride bytes, 16-aligned): [0, HDR_RESERVE) header, [HDR_RESERVE, ...) payload words.
// ------------------------------------------------------------------------------------------
struct EncArgs {
    const uint8_t *src;
    size_t n;
    uint32_t block_size, nblocks;
    uint32_t req_log2, n_states, tlmax;
    const uint32_t *counts;  // [nblocks*256] (per-block mode)
    uint8_t *scratch;
    size_t stride;
    uint32_t pay_cap_words;
    uint32_t *hlen, *plen;   // [nblocks] header / payload bytes
    int32_t *status;
    GlobalTable g;
    int global_mode;
    // segmented per-block mode (fse_shared_enc.cuh): a block = segs_per_block streams of seg_size bytes sharing the block's
    // table; scratch slots, hlen, plen and status are per stream
    uint32_t seg_size, segs_per_block;
    uint32_t flags;          // FSE_B200_FLAG_* (per-block tables, one stream per block)
};

// FSE_B200_FLAG_RAW_IF_EXPANDS: a block whose header + payload would not be smaller than 1 + its length is stored as
// 0x0F + raw bytes.  The reference has no such fallback (fse.rs:191-193 only bounds the expansion), so
// this only runs when the caller asks for it.  Returns true when the block was replaced (hl = 1, pl = bn, status 1).
__device__ __forceinline__ bool warp_raw_if_expands(uint32_t flags, uint32_t hl, uint32_t pl, uint8_t *bs, const uint8_t *__restrict__ bsrc,
                                                    uint32_t bn, int lane)
{
    if (!(flags & 1u) || hl + pl < 1 + bn) return false;
    for (uint32_t i = lane; i < bn; i += 32) bs[512 + i] = bsrc[i];      // the payload area starts at HDR_RESERVE
    if (lane == 0) bs[0] = 0x0F;
    return true;
}

// fse.rs:210-218
__device__ __forceinline__ uint32_t enc_first(const uint16_t *tab, const uint2 *tt, uint32_t sym)
{
    uint2 t = tt[sym];
    uint32_t bo = (t.x + (1u << 15)) >> 16;
    uint32_t value = (bo << 16) - t.x;
    return tab[(int32_t)(value >> bo) + (int32_t)t.y];
}

__device__ void encode_payload_warp(const uint8_t *__restrict__ bsrc, uint32_t bn, uint32_t N, uint32_t log2,
                                    const uint16_t *tab, const uint2 *tt, uint32_t *fld, uint32_t *rows,
                                    uint32_t *pay, uint32_t cap_words, int lane, uint32_t &bits_out, bool &overflow)
{
    const bool act = (uint32_t)lane < N;
    const uint32_t c = (bn - 1) & (N - 1);
    const uint32_t kcol = act ? ((c - (uint32_t)lane) & (N - 1)) : (uint32_t)lane;  // stream position in a round
    uint32_t state = 0;
    int32_t i0 = (int32_t)bn - 1 - (int32_t)kcol;  // my highest symbol (costs no bits, lib.rs:123,155-165)
    if (act) state = enc_first(tab, tt, __ldg(bsrc + i0));
    i0 -= (int32_t)N;
    const uint32_t G = (bn - N + N - 1) / N;  // rounds of N transitions
    uint32_t cw = 0, cb = 0, wdone = 0;
    uint32_t *myrow = rows + lane * ROW_STRIDE;
    uint32_t *obuf = fld;                      // fld is dead between pass 2 and the next pass 1
    overflow = false;

    // my 32 symbols of a chunk, fetched one chunk ahead so that DRAM latency hides behind pass 2
    uint32_t sy[32];
    const uint32_t NOSYM = 0x100;
#pragma unroll
    for (int r = 0; r < 32; r++) {
        int32_t ii = i0 - (int32_t)(r * N);
        sy[r] = (act && ii >= 0) ? (uint32_t)__ldg(bsrc + ii) : NOSYM;
    }
    for (uint32_t g0 = 0; g0 < G; g0 += 32) {
        // pass 1: 32 rounds of the state transform (fse.rs:227-239); fields go to fld in stream order,
        // row r, 16-byte chunk (kcol>>2) ^ (r&7) (conflict free for the writer and for pass 2's reader)
#pragma unroll
        for (int r = 0; r < 32; r++) {
            uint32_t f = 0;
            if (sy[r] != NOSYM) {
                uint2 t = tt[sy[r]];
                uint32_t bo = (t.x + state) >> 16;
                f = (state & ((1u << bo) - 1u)) | (bo << 16);
                state = tab[(int32_t)(state >> bo) + (int32_t)t.y];
            }
            fld[r * 32 + ((((kcol >> 2) ^ (r & 7)) << 2) | (kcol & 3))] = f;
        }
        {   // prefetch the next chunk's symbols
            int32_t base = i0 - (int32_t)((g0 + 32) * N);
#pragma unroll
            for (int r = 0; r < 32; r++) {
                int32_t ii = base - (int32_t)(r * N);
                sy[r] = (act && ii >= 0) ? (uint32_t)__ldg(bsrc + ii) : NOSYM;
            }
        }
        __syncwarp();
        // pass 2: lane L serialises round L (32 consecutive fields of the stream)
        BitRow br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint4 x = *reinterpret_cast<const uint4 *>(fld + lane * 32 + ((q ^ (lane & 7)) << 2));
            br.put(x.x & 0xffff, x.x >> 16);
            br.put(x.y & 0xffff, x.y >> 16);
            br.put(x.z & 0xffff, x.z >> 16);
            br.put(x.w & 0xffff, x.w >> 16);
        }
        uint32_t tot = br.finish();
        __syncwarp();
        // concatenate into obuf, then stream the full words out as whole 128-byte lines
        uint32_t nw = warp_place(myrow, tot, obuf, 1024, lane, cw, cb, overflow);
        __syncwarp();
        if (wdone + nw > cap_words) { overflow = true; nw = 0; }
        for (uint32_t j = lane; j < nw; j += 32) pay[wdone + j] = obuf[j];
        wdone += nw;
        __syncwarp();
    }
    // final states N-1 .. 0 (fse.rs:248-250, order lib.rs:178-179), then the marker bit (lib.rs:141,181)
    {
        uint32_t st = __shfl_sync(FULL, state, (N - 1 - lane) & 31);
        BitRow br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
        if (act) br.put(st & ((1u << log2) - 1u), log2);
        if ((uint32_t)lane == N - 1) br.put(1, 1);
        uint32_t tot = br.finish();
        __syncwarp();
        wdone += warp_place(myrow, tot, pay + wdone, cap_words > wdone ? cap_words - wdone : 0, lane, cw, cb, overflow);
        __syncwarp();
    }
    if (cb) {
        if (wdone < cap_words) { if (lane == 0) pay[wdone] = cw; }
        else overflow = true;
    }
    bits_out = wdone * 32 + cb;
}

__global__ void __launch_bounds__(512) k_encode_blocks(EncArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const EncLayout lay = enc_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint16_t *tab = reinterpret_cast<uint16_t *>(my + lay.tab);
    uint2 *tt = reinterpret_cast<uint2 *>(my + lay.tt);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(my + lay.work);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.work + 1024);
    uint32_t *cum = reinterpret_cast<uint32_t *>(my + lay.work + 2048);
    uint8_t *spread = my + lay.work + 3072;
    uint32_t *fld = reinterpret_cast<uint32_t *>(my + lay.work);
    uint32_t *rows = reinterpret_cast<uint32_t *>(my + lay.rows);
    const uint32_t N = a.n_states;

    uint32_t glog2 = 0;
    if (a.global_mode) {  // the shared table is loaded once per warp
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) tab[i] = a.g.enc_table[i];
        for (uint32_t i = lane; i < 256; i += 32) tt[i] = a.g.enc_tt[i];
        __syncwarp();
    }

    for (uint32_t b = blockIdx.x * wpc + warp; b < a.nblocks; b += gridDim.x * wpc) {
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        const uint8_t *bsrc = a.src + off;
        uint8_t *bs = a.scratch + (size_t)b * a.stride;
        uint32_t *hdr_words = reinterpret_cast<uint32_t *>(bs);
        uint32_t *pay = reinterpret_cast<uint32_t *>(bs + HDR_RESERVE);
        uint32_t log2 = glog2, hbytes = 0;
        int st = ST_OK;

        if (!a.global_mode) {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; k++) cnt[k * 32 + lane] = a.counts[(size_t)b * 256 + k * 32 + lane];
            __syncwarp();
            uint32_t table_len;
            int rc = warp_normalize(cnt, (uint64_t)bn, a.req_log2, norm, lane, log2, table_len);
            if (rc < 0) {
                // blocks the reference panics on: stored with an escape byte (include/fse_b200.h)
                if (table_len <= 1) {            // all bytes zero: histogram.rs:98
                    if (lane == 0) { bs[0] = 0x0E; bs[1] = 0x00; a.hlen[b] = 2; a.plen[b] = 0; a.status[b] = 2; }
                } else if (bn <= 4) {            // histogram.rs:271
                    if (lane == 0) {
                        bs[0] = 0x0F;
                        for (uint32_t i = 0; i < bn; i++) bs[1 + i] = bsrc[i];
                        a.hlen[b] = 1 + bn; a.plen[b] = 0; a.status[b] = 1;
                    }
                } else if (lane == 0) { a.hlen[b] = 0; a.plen[b] = 0; a.status[b] = rc; }
                continue;
            }
            if (bn < N) {                        // fewer symbols than states: lib.rs:121,154,156
                if ((uint32_t)lane < bn) bs[1 + lane] = bsrc[lane];
                if (lane == 0) { bs[0] = 0x0F; a.hlen[b] = 1 + bn; a.plen[b] = 0; a.status[b] = 1; }
                continue;
            }
            if (log2 > a.tlmax) {
                if (lane == 0) { a.hlen[b] = 0; a.plen[b] = 0; a.status[b] = ST_UNSUPPORTED; }
                continue;
            }
            uint32_t hbits = warp_ncount_write(norm, log2, table_len, rows, hdr_words, lane);
            hbytes = (hbits + 7) >> 3;
            warp_spread(norm, log2, table_len, spread, cum, tab, lane);
            warp_build_encode(norm, log2, table_len, spread, cum, tab, tt, lane);
        } else if (bn < N) {                     // global mode: a short tail is stored raw, no escape
            if ((uint32_t)lane < bn) bs[lane] = bsrc[lane];
            if (lane == 0) { a.hlen[b] = bn; a.plen[b] = 0; a.status[b] = 1; }
            continue;
        }
        uint32_t pbits;
        bool ovf;
        encode_payload_warp(bsrc, bn, N, log2, tab, tt, fld, rows, pay, a.pay_cap_words, lane, pbits, ovf);
        if (ovf) st = ST_CAPACITY;
        uint32_t hl = ovf ? 0 : hbytes, pl = ovf ? 0 : (pbits + 7) >> 3;
        if (!ovf && !a.global_mode && warp_raw_if_expands(a.flags, hl, pl, bs, bsrc, bn, lane)) { hl = 1; pl = bn; st = 1; }
        if (lane == 0) {
            a.hlen[b] = hl;
            a.plen[b] = pl;
            a.status[b] = st;
        }
    }
}

// exclusive scan of block sizes -> uint64 offsets[nblocks+1]; single CTA of 1024 threads, tiles of 8 192 blocks.
// The sizes of a tile are loaded coalesced into shared memory (padded: one word per 32, so that a thread's eight
// consecutive entries are conflict free), every thread scans its eight, one block scan gives the bases.  With many
// small blocks the first version (each thread walking its own slice of global memory) was 15 % of the compress
// time (65 536 blocks of 4 KiB: 0.19 ms).
constexpr int SCAN_E = 8, SCAN_TILE = 1024 * SCAN_E;
__global__ void __launch_bounds__(1024) k_scan_sizes(const uint32_t *__restrict__ hlen, const uint32_t *__restrict__ plen,
                                                      uint32_t nblocks, unsigned long long *offsets)
{
    __shared__ uint32_t sz[SCAN_TILE + SCAN_TILE / 32];
    __shared__ unsigned long long wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long carry = 0;
    for (uint32_t t0 = 0; t0 < nblocks; t0 += SCAN_TILE) {
#pragma unroll
        for (int k = 0; k < SCAN_E; k++) {
            const uint32_t e = k * 1024 + tid, b = t0 + e;
            sz[e + (e >> 5)] = b < nblocks ? hlen[b] + plen[b] : 0u;
        }
        __syncthreads();
        uint32_t v[SCAN_E];
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_E; k++) {
            const uint32_t e = tid * SCAN_E + k;
            v[k] = sz[e + (e >> 5)];
            s += v[k];
        }
        unsigned long long incl = warp_incl_add(s, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane];
            unsigned long long wi = warp_incl_add(w, lane);
            wsum[lane] = wi - w;
        }
        __syncthreads();
        unsigned long long run = carry + wsum[warp] + incl - s;
        const uint32_t b0 = t0 + tid * SCAN_E;
#pragma unroll
        for (int k = 0; k < SCAN_E; k++) {
            if (b0 + k < nblocks) offsets[b0 + k] = run;
            run += v[k];
        }
        // the tile total: the last thread's running sum
        __syncthreads();
        if (tid == 1023) wsum[0] = run;
        __syncthreads();
        carry = wsum[0];
        __syncthreads();
    }
    if (tid == 0) offsets[nblocks] = carry;
}

#ifndef GATHER_UNROLL
#define GATHER_UNROLL 2
#endif
// byte copy with 4-byte aligned destination stores and funnel-shifted source words
__device__ __forceinline__ void block_copy_bytes(uint8_t *dst, const uint8_t *src /*4-aligned*/, uint32_t len, int tid, int nthr)
{
    // destination stores are whole aligned 16-byte vectors; the source words are funnel-shifted into place
    uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > len) head = len;
    if (tid < (int)head) dst[tid] = src[tid];
    const uint32_t nvec = (len - head) >> 4;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src) + (head >> 2);   // vector k = src bytes head+16k ..
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    const uint32_t sh = (head & 3) * 8;
    // GATHER_UNROLL vectors per thread in flight: the copy is bound by the bytes it keeps in flight, not by instructions
    uint32_t k = tid;
    for (; k + (GATHER_UNROLL - 1) * nthr < nvec; k += GATHER_UNROLL * nthr) {
        uint32_t x[GATHER_UNROLL][5];
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; u++) {
            const uint32_t *w = sw + 4 * (k + u * nthr);
            x[u][0] = w[0]; x[u][1] = w[1]; x[u][2] = w[2]; x[u][3] = w[3]; x[u][4] = sh ? w[4] : 0u;
        }
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; u++)
            dv[k + u * nthr] = make_uint4(__funnelshift_r(x[u][0], x[u][1], sh), __funnelshift_r(x[u][1], x[u][2], sh),
                                          __funnelshift_r(x[u][2], x[u][3], sh), __funnelshift_r(x[u][3], x[u][4], sh));
    }
    for (; k < nvec; k += nthr) {
        const uint32_t *w = sw + 4 * k;
        uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
        dv[k] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                           __funnelshift_r(w3, w4, sh));
    }
    const uint32_t t0 = head + (nvec << 4);
    if (t0 + tid < len) dst[t0 + tid] = src[t0 + tid];    // tail < 16 bytes <= nthr
}

// K6: gather header || payload of every block to its scanned offset
__global__ void __launch_bounds__(256) k_gather(const uint8_t *__restrict__ scratch, size_t stride,
                                                 const uint32_t *__restrict__ hlen, const uint32_t *__restrict__ plen,
                                                 const unsigned long long *__restrict__ offsets, uint32_t nblocks,
                                                 uint8_t *__restrict__ dst)
{
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const uint8_t *bs = scratch + (size_t)b * stride;
        uint8_t *d = dst + offsets[b];
        uint32_t h = hlen[b], p = plen[b];
        block_copy_bytes(d, bs, h, threadIdx.x, blockDim.x);
        block_copy_bytes(d + h, bs + HDR_RESERVE, p, threadIdx.x, blockDim.x);
    }
}

// ------------------------------------------------------------------------------------------
// K5: per block  header parse -> decode table -> stack-order N-state decode, length driven.
// ------------------------------------------------------------------------------------------
struct DecArgs {
    const uint8_t *comp;
    size_t comp_bytes;
    const unsigned long long *offsets;
    uint32_t nblocks, block_size;
    size_t n;
    uint32_t n_states, tlmax;
    uint8_t *dst;
    int32_t *status;
    GlobalTable g;
    int global_mode;
    uint32_t seg_size, segs_per_block;   // segmented per-block mode: offsets / status are per stream (fse_shared_dec.cuh)
    uint32_t dec_copies;                 // copies of the CTA-owned decode table (32 / 16 / 8)
    // exhaust mode (the reference's own termination rule, src/lib.rs:198,228: decode until the bit
    // stack cannot supply num_bits): block b may produce up to block_size bytes, out_len[b] = produced
    int exhaust;
    uint32_t *out_len;
};

// What every block decoder does before it looks at a header: the offset checks, the short raw tail of the global-table
// mode and the escape blocks (include/fse_b200.h: 0x0F raw, 0x0E run; no valid header starts with them, histogram.rs:439-441).
// Returns true when the block is finished (st = its status word); otherwise cs / clen are the stream to parse.  All lanes call.
__device__ __forceinline__ bool dec_block_prologue(const DecArgs &a, uint32_t b, uint32_t bn, uint32_t N, uint8_t *out, int lane,
                                                   const uint8_t *&cs, uint32_t &clen, int &st)
{
    const unsigned long long o0 = a.offsets[b], o1 = a.offsets[b + 1];
    if (o1 < o0 || o1 > a.comp_bytes || o1 - o0 > 0xffffffffull) { st = ST_LENGTH; return true; }
    cs = a.comp + o0;
    clen = (uint32_t)(o1 - o0);
    if (a.global_mode) {
        if (bn >= N) return false;
        if (clen != bn) { st = ST_LENGTH; return true; }     // a short tail is stored raw, no escape
        for (uint32_t i = lane; i < bn; i += 32) out[i] = cs[i];
        st = 1;
        return true;
    }
    if (clen == 0) { st = ST_PANIC; return true; }           // stream_reader.rs:17
    const uint32_t first = cs[0];
    if (a.exhaust && (first & 0x0f) > 10) { st = ST_TABLE_LOG; return true; }   // TableLogTooLarge -> None, histogram.rs:439-441, lib.rs:191,219
    if ((first & 0x0f) == 0x0f) {                            // raw escape
        if (clen != 1 + bn) { st = ST_LENGTH; return true; }
        for (uint32_t i = lane; i < bn; i += 32) out[i] = cs[1 + i];
        st = 1;
        return true;
    }
    if ((first & 0x0f) == 0x0e) {                            // run escape
        if (clen != 2) { st = ST_LENGTH; return true; }
        const uint8_t v = cs[1];
        for (uint32_t i = lane; i < bn; i += 32) out[i] = v;
        st = 2;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(512) k_decode_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const DecLayout lay = dec_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint32_t *tab = reinterpret_cast<uint32_t *>(my + lay.tab);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.norm);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.ctr);
    uint8_t *spread = my + lay.spread;
    uint32_t *ring = reinterpret_cast<uint32_t *>(my + lay.ring);
    const uint32_t N = a.n_states;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) tab[i] = a.g.dec_table[i];
        __syncwarp();
    }

    for (uint32_t b = blockIdx.x * wpc + warp; b < a.nblocks; b += gridDim.x * wpc) {
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        uint8_t *out = a.dst + off;
        int st = ST_OK;
        const uint8_t *cs;
        uint32_t clen;
        if (dec_block_prologue(a, b, bn, N, out, lane, cs, clen, st)) {      // bad offsets, raw tail, escape blocks
            if (lane == 0) { a.status[b] = st; if (a.exhaust) a.out_len[b] = 0; }
            continue;
        }
        uint32_t log2 = glog2, consumed = 0;
        __syncwarp();
        if (!a.global_mode) {
#pragma unroll
            for (int k = 0; k < 8; k++) norm[k * 32 + lane] = 0;
            __syncwarp();
            uint32_t table_len = 0;
            int rc = 0;
            if (lane == 0) rc = ncount_read_serial(cs, clen, norm, log2, table_len, consumed);
            rc = __shfl_sync(FULL, rc, 0);
            log2 = __shfl_sync(FULL, log2, 0);
            table_len = __shfl_sync(FULL, table_len, 0);
            consumed = __shfl_sync(FULL, consumed, 0);
            __syncwarp();
            if (rc < 0) { if (lane == 0) a.status[b] = rc; continue; }
            if (log2 > a.tlmax) { if (lane == 0) a.status[b] = ST_UNSUPPORTED; continue; }
            warp_spread(norm, log2, table_len, spread, ctr, reinterpret_cast<uint16_t *>(tab), lane);
            warp_build_decode(norm, log2, table_len, spread, ctr, tab, lane);
        }
        if (bn < N) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        // BitStackReader::new, stack_reader.rs:17-92: the highest set bit of the last byte is the marker
        const uint8_t *pay = cs + consumed;
        const uint32_t plen = clen - consumed;
        if (plen == 0 || pay[plen - 1] == 0) { if (lane == 0) a.status[b] = ST_NO_MARKER; continue; }
        // The payload is staged through a 256-word ring in shared memory, refilled 128 words at a time
        // from registers that were loaded one refill earlier (DRAM latency never sits in the state chain).
        // Bit positions below are relative to `origin`, the 4-byte aligned word holding the first byte.
        const uint32_t bias = (uint32_t)((uintptr_t)pay & 3);
        const uint32_t *origin = reinterpret_cast<const uint32_t *>(pay - bias);
        uint32_t cur = (plen - 1) * 8 + ilog2u(pay[plen - 1]) + 8 * bias;   // marker position
        const uint32_t floor_bits = 8 * bias;                                 // stream bit 0
        if (cur - floor_bits < N * log2) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }   // lib.rs:197,224-225
        const uint32_t topq = cur >> 5;
        uint32_t lowq = (topq & ~127u) >= 128 ? (topq & ~127u) - 128 : 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t w = lowq + lane + 32 * k;
            if (w <= topq) ring[w & 255] = __ldg(origin + w);
        }
        uint32_t pre[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
        __syncwarp();
        auto ring_bits = [&](uint32_t q, uint32_t nb) -> uint32_t {
            uint32_t w = q >> 5;
            return __funnelshift_r(ring[w & 255], ring[(w + 1) & 255], q & 31) & ((1u << nb) - 1u);
        };
        const bool act = (uint32_t)lane < N;
        // Decoder::new, fse.rs:349-352: state 0 is read first (it was written last)
        uint32_t state = act ? ring_bits(cur - (lane + 1) * log2, log2) : 0u;
        cur -= N * log2;
        const uint32_t body = bn - N;   // exhaust mode: bn == capacity
        bool bad = false;
        uint32_t stop_i = body, stop_lane = body & (N - 1);
        for (uint32_t i0 = 0;; i0 += N) {
            if (i0 >= body) {
                if (a.exhaust) bad = true;                 // ran past the capacity (quirk Q1)
                break;
            }
            if ((cur >> 5) < lowq + 17 && lowq) {          // a round takes at most 15 words: refill the ring
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; k++) ring[(lowq - 128 + lane + 32 * k) & 255] = pre[k];
                lowq -= 128;
#pragma unroll
                for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
                __syncwarp();
            }
            uint32_t i = i0 + lane;
            bool on = act && i < body;
            uint32_t e = tab[state];                       // fse.rs:363-373
            uint32_t nb = on ? (e >> 24) : 0u;
            uint32_t incl = warp_incl_add(nb, lane);
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) {                  // the stack cannot supply this round
                if (!a.exhaust) { bad = true; break; }
                // reference rule: the first decoder that gets None stops the loop (lib.rs:198,228-241)
                uint32_t okm = __ballot_sync(FULL, on && incl <= cur - floor_bits);
                uint32_t s = __popc(okm);                  // lanes [0, s) still decode
                if (on && (uint32_t)lane < s) {
                    uint32_t bits = ring_bits(cur - incl, nb);
                    out[i] = (uint8_t)(e >> 16);
                    state = (e & 0xffffu) + bits;
                }
                uint32_t used = s ? __shfl_sync(FULL, incl, (s - 1) & 31) : 0u;
                cur -= used;
                if (i0 + s >= body) bad = true;            // would not fit the capacity
                stop_i = i0 + s;
                stop_lane = s & (N - 1);
                break;
            }
            if (on) {
                uint32_t bits = ring_bits(cur - incl, nb);
                out[i] = (uint8_t)(e >> 16);
                state = (e & 0xffffu) + bits;
            }
            cur -= tot;
        }
        cur -= floor_bits;
        if (!bad && act) {                                 // Decoder::finish, fse.rs:383-385 (order lib.rs:236-243)
            uint32_t i = stop_i + ((lane - stop_lane) & (N - 1));
            out[i] = (uint8_t)(tab[state] >> 16);
        }
        if (a.exhaust) {
            if (lane == 0) { a.out_len[b] = bad ? 0 : stop_i + N; a.status[b] = bad ? ST_CAPACITY : ST_OK; }
            continue;
        }
        if (bad || cur != 0) st = ST_LENGTH;
        if (lane == 0) a.status[b] = st;
    }
}

// ------------------------------------------------------------------------------------------
// Stage kernels (one warp per table) behind the stage entry points of the C ABI
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_normalize(const unsigned long long *__restrict__ counts64, uint32_t ntables,
                                                   uint32_t req_log2, int32_t *norm_out, uint32_t *log2_out,
                                                   uint32_t *table_len_out, int32_t *status)
{
    __shared__ int32_t norm[256];
    const int lane = threadIdx.x;
    for (uint32_t t = blockIdx.x; t < ntables; t += gridDim.x) {
        const unsigned long long *c = counts64 + (size_t)t * 256;
        unsigned long long s = 0;
        for (int k = 0; k < 8; k++) s += c[lane * 8 + k];
#pragma unroll
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(FULL, s, d);
        uint32_t log2 = 0, table_len = 0;
        __syncwarp();
        int rc = warp_normalize(c, (uint64_t)s, req_log2, norm, lane, log2, table_len);
        for (int k = 0; k < 8; k++) norm_out[(size_t)t * 256 + lane * 8 + k] = rc < 0 ? 0 : norm[lane * 8 + k];
        if (lane == 0) { log2_out[t] = log2; table_len_out[t] = table_len; status[t] = rc; }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) k_ncount_write(const int32_t *__restrict__ norm_in, const uint32_t *__restrict__ log2,
                                                      const uint32_t *__restrict__ table_len, uint32_t ntables,
                                                      uint8_t *out, size_t stride, uint32_t *bytes, uint32_t *bits)
{
    __shared__ int32_t norm[256];
    __shared__ uint32_t rows[ROWS_WORDS];
    __shared__ __align__(16) uint32_t hdr[HDR_RESERVE / 4];
    const int lane = threadIdx.x;
    for (uint32_t t = blockIdx.x; t < ntables; t += gridDim.x) {
        for (int k = 0; k < 8; k++) norm[k * 32 + lane] = norm_in[(size_t)t * 256 + k * 32 + lane];
        __syncwarp();
        uint32_t hb = warp_ncount_write(norm, log2[t], table_len[t], rows, hdr, lane);
        uint32_t nbytes = (hb + 7) >> 3;
        __syncwarp();
        const uint8_t *h8 = reinterpret_cast<const uint8_t *>(hdr);
        for (uint32_t i = lane; i < nbytes && i < stride; i += 32) out[(size_t)t * stride + i] = h8[i];
        if (lane == 0) { bytes[t] = nbytes; bits[t] = hb; }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) k_ncount_read(const uint8_t *__restrict__ in, size_t stride, const uint32_t *__restrict__ len,
                                                     uint32_t ntables, int32_t *norm_out, uint32_t *log2_out,
                                                     uint32_t *table_len_out, uint32_t *consumed_out, int32_t *status)
{
    __shared__ int32_t norm[256];
    const int lane = threadIdx.x;
    for (uint32_t t = blockIdx.x; t < ntables; t += gridDim.x) {
        for (int k = 0; k < 8; k++) norm[k * 32 + lane] = 0;
        __syncwarp();
        uint32_t log2 = 0, table_len = 0, consumed = 0;
        int rc = 0;
        if (lane == 0) rc = ncount_read_serial(in + (size_t)t * stride, len[t], norm, log2, table_len, consumed);
        __syncwarp();
        for (int k = 0; k < 8; k++) norm_out[(size_t)t * 256 + k * 32 + lane] = norm[k * 32 + lane];
        if (lane == 0) { log2_out[t] = log2; table_len_out[t] = table_len; consumed_out[t] = consumed; status[t] = rc; }
        __syncwarp();
    }
}

// dynamic smem: norm i32[256] | cum u32[256] | spread u8[size] | table (u16 or u32)[size] | tt uint2[256]
__global__ void __launch_bounds__(32) k_build_tables(const int32_t *__restrict__ norm_in, const uint32_t *__restrict__ log2_in,
                                                      const uint32_t *__restrict__ table_len_in, uint32_t ntables,
                                                      uint32_t max_log2, int decode, uint16_t *enc_table, uint2 *enc_tt,
                                                      uint8_t *symbols, uint32_t *dec_table, int32_t *status)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t msize = 1u << max_log2;
    int32_t *norm = reinterpret_cast<int32_t *>(smem_raw);
    uint32_t *cum = reinterpret_cast<uint32_t *>(smem_raw + 1024);
    uint2 *tt = reinterpret_cast<uint2 *>(smem_raw + 2048);
    uint32_t *table32 = reinterpret_cast<uint32_t *>(smem_raw + 4096);
    uint16_t *table16 = reinterpret_cast<uint16_t *>(smem_raw + 4096);
    uint8_t *spread = smem_raw + 4096 + (size_t)msize * 4;
    const int lane = threadIdx.x;
    for (uint32_t t = blockIdx.x; t < ntables; t += gridDim.x) {
        const uint32_t log2 = log2_in[t], table_len = table_len_in[t];
        if (log2 < TL_MIN || log2 > max_log2 || table_len == 0 || table_len > 256) {
            if (lane == 0) status[t] = log2 > max_log2 ? ST_UNSUPPORTED : ST_PANIC;
            continue;
        }
        const uint32_t size = 1u << log2;
        for (int k = 0; k < 8; k++) norm[k * 32 + lane] = norm_in[(size_t)t * 256 + k * 32 + lane];
        __syncwarp();
        warp_spread(norm, log2, table_len, spread, cum, table16, lane);
        if (!decode) {
            warp_build_encode(norm, log2, table_len, spread, cum, table16, tt, lane);
            for (uint32_t i = lane; i < size; i += 32) enc_table[(size_t)t * msize + i] = table16[i];
            for (uint32_t i = lane; i < 256; i += 32) enc_tt[(size_t)t * 256 + i] = tt[i];
            if (symbols) for (uint32_t i = lane; i < size; i += 32) symbols[(size_t)t * msize + i] = spread[i];
        } else {
            warp_build_decode(norm, log2, table_len, spread, cum, table32, lane);
            for (uint32_t i = lane; i < size; i += 32) dec_table[(size_t)t * msize + i] = table32[i];
        }
        if (lane == 0) status[t] = 0;
        __syncwarp();
    }
}

// Global-table mode: a symbol the installed table gives no probability must not be able to take a coder outside its
// tables (its reference transform, fse.rs:170, indexes past the next-state table, and the entries of symbols >= table_len
// are zero).  Such symbols get a transform with num_bits = table_log and next state = table[0]: what a block that
// contains one encodes to cannot be decoded (include/fse_b200.h says so, like the crate), but nothing can fault.
__global__ void k_sanitize_global_tt(uint2 *tt, const int32_t *__restrict__ norm, uint32_t log2)
{
    const uint32_t i = threadIdx.x;
    if (i < 256 && norm[i] == 0) tt[i] = make_uint2((log2 << 16) - (1u << log2), 0xffffffffu);
}

// widen uint32 counts to uint64 (stage API / global table plumbing)
__global__ void k_widen_counts(const uint32_t *__restrict__ in, unsigned long long *out, size_t count)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = in[i];
}

// synthetic generators of SURVEY.md 8(d): byte i = LUT[r16(seed, i) & (lut_len-1)]
__global__ void k_generate(const uint8_t *__restrict__ lut, uint32_t lut_mask, uint64_t seed, uint64_t first_index,
                           uint8_t *__restrict__ dst, size_t n)
{
    // one thread per group of four bytes that share one splitmix64 draw
    size_t q0 = first_index >> 2;
    size_t nq = ((first_index + n + 3) >> 2) - q0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = splitmix64(seed + q0 + q);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t i = ((q0 + q) << 2) + k;
            if (i >= first_index && i < first_index + n)
                dst[i - first_index] = lut[(uint32_t)((z >> (16 * k)) & 0xFFFF) & lut_mask];
        }
    }
}

}  // namespace fsed
