// fse_shared_enc.cuh -- encode with tables OWNED BY THE CTA and replicated across the 32 shared-memory banks.
//
// The private-table kernels (fse_encode128.cuh) pay 2.2-3.4 shared-memory wavefronts for each of the two random
// table look-ups of a tANS step (fse.rs:227-239): 32 lanes hit 32 banks at random.  Here one table set serves every
// warp of the CTA and is stored once per bank: lane l reads copy l, which lives entirely in bank l, so every look-up
// is exactly one wavefront whatever the states are.
//   symbol transforms  ttR : uint32 P[sym] at  ttR + sym * 128 + lane * 4                       (32 KiB)
//   next-state table   tabR: uint16 pairs, entries 2j / 2j+1 in the word at tabR + j * 128 + lane * 4   (size * 64 B)
// 160 KiB at table_log 11, which is why the table must be shared by the CTA: one table in flight per SM.  Users:
// global-table mode (one table per job) and the segmented per-block mode (one table per block, the block's segments
// are coded by the warps of one CTA).
//
// The transform is re-packed so that a step needs few ALU-pipe instructions (the ALU pipe issues one warp instruction
// per two cycles and is the next limiter once the bank conflicts are gone):
//   P = H << 12 | (find_state + 2048),   H = (max_bits << 13) - (count << max_bits)
//   t = P + (state << 12)                 one IMAD (FMA pipe)
//   nb = t >> 25                          = (bits + state) >> 16 of fse.rs:228: state - (count << max_bits) lies in
//                                           (-2^12, 2^12), so a 13-bit fraction decides max_bits vs max_bits - 1
//   u = (t & 0xfff) + (state >> nb)       = find_state + 2048 + (state >> nb): the table index, biased by 2048
// and the emitted bits are never masked out of the state: a funnel shift moves the low nb bits of the raw state into
// the top of the pair accumulator.  table_log <= 11.
#pragma once
#include "fse_encode128.cuh"

namespace fsed {

constexpr uint32_t SH_TL_MAX = 11;
// sh_pack_tt (reference transform -> P), SH_FS_BIAS and quad_field_raw live in fse_encode128.cuh: the private-table
// kernel uses the same arithmetic.

// NSR = copies of the next-state table: 32 (lane l reads copy l, one wavefront, size * 64 bytes) or 16 (lanes l and
// l + 16 share copy l & 15 in banks l & 15 / (l & 15) + 16, at most two wavefronts, size * 32 bytes: room for twice the warps)
struct ShEnc { uint32_t ttl, tabl; };     // ttR + lane * 4;  tabR - SH_FS_BIAS * 2 * NSR + (lane & (NSR - 1)) * 4

// shared address of the lane's copy of P[byte k of x]: PRMT (ALU pipe) + IMAD (FMA pipe)
template <int K>
__device__ __forceinline__ uint32_t sh_tt_addr(const ShEnc &e, uint32_t x)
{
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(a) : "r"(__byte_perm(x, 0u, 0x4440u + K)), "r"(e.ttl));
    return a;
}
template <int NSR>
__device__ __forceinline__ uint32_t sh_tab_addr(const ShEnc &e, uint32_t u)
{
    // (u >> 1) * 4 * NSR + (u & 1) * 2 = u * 2 * NSR - (2 * NSR - 2) * (u & 1): one LOP on the ALU pipe, two IMADs on the
    // FMA pipe (written as mad.lo so that the compiler does not turn the second product into compare + select)
    uint32_t a;
    if (NSR == 32) {
        asm("mad.lo.u32 %0, %1, 64, %2;" : "=r"(a) : "r"(u), "r"(e.tabl));
        asm("mad.lo.u32 %0, %1, 0xffffffc2, %0;" : "+r"(a) : "r"(u & 1u));
    } else {
        asm("mad.lo.u32 %0, %1, 32, %2;" : "=r"(a) : "r"(u), "r"(e.tabl));
        asm("mad.lo.u32 %0, %1, 0xffffffe2, %0;" : "+r"(a) : "r"(u & 1u));
    }
    return a;
}
// one transition (fse.rs:227-239): the caller emits the low nb bits of the OLD state
template <int NSR>
__device__ __forceinline__ uint32_t sh_enc_step_at(const ShEnc &e, uint32_t tt_addr, uint32_t s, uint32_t &nb)
{
    const uint32_t t = lds_u32(tt_addr) + (s << 12);
    nb = t >> 25;
    const uint32_t u = (t & 0xfffu) + (s >> nb);
    return lds_u16(sh_tab_addr<NSR>(e, u));
}
template <int NSR>
__device__ __forceinline__ uint32_t sh_enc_step(const ShEnc &e, uint32_t sym, uint32_t s, uint32_t &nb)
{
    return sh_enc_step_at<NSR>(e, e.ttl + (sym << 7), s, nb);
}
// Encoder::new_first_symbol (fse.rs:210-218): bo = max_bits, value = count << max_bits, so the state is table[total]
template <int NSR>
__device__ __forceinline__ uint32_t sh_enc_first(const ShEnc &e, uint32_t sym)
{
    const uint32_t p = lds_u32(e.ttl + (sym << 7));
    const uint32_t H = p >> 12;
    const uint32_t mbo = (H + 8191u) >> 13;
    const uint32_t x = ((mbo << 13) - H) >> mbo;
    return lds_u16(sh_tab_addr<NSR>(e, (p & 0xfffu) + x));
}

__device__ __forceinline__ uint2 sh_quad_field(uint32_t s3, uint32_t b3, uint32_t s2, uint32_t b2, uint32_t s1, uint32_t b1,
                                               uint32_t s0, uint32_t b0)
{
    return quad_field_raw(s3, b3, s2, b2, s1, b1, s0, b0);
}

template <int ROUNDS> struct ShEncStage {
    static constexpr int FLD_WORDS = ROUNDS * 64;                             // ROUNDS rows of 32 quad fields
    // a lane string: ROUNDS quads of at most 44 bits (table_log <= 11) + 31 carried bits, + the word finish() stores
    static constexpr int ROW_STRIDE = (ROUNDS == 16) ? 25 : 15;               // words per lane string, odd
    static constexpr int ROWS_WORDS = 32 * ROW_STRIDE;
    static constexpr int BYTES = (FLD_WORDS + ROWS_WORDS) * 4;
};

template <int NSR>
__device__ __forceinline__ void sh_elem_checked(const uint8_t *__restrict__ bsrc, int32_t i, int32_t bn, const ShEnc &e,
                                                uint32_t &s, uint32_t &sold, uint32_t &nb)
{
    sold = 0; nb = 0;
    if (i < 0 || i >= bn) return;
    const uint32_t sym = __ldg(bsrc + i);
    if (i >= bn - 128) s = sh_enc_first<NSR>(e, sym);
    else { sold = s; s = sh_enc_step<NSR>(e, sym, s, nb); }
}

// 128-state payload of one stream (block or segment) by one warp; the structure of encode128_payload_warp with the
// shared tables.  ROUNDS = rounds per chunk (16: lane L serialises half a round of 32 quads; 8: a quarter).
template <int ROUNDS, int NSR>
__device__ void sh_encode_payload_warp(const uint8_t *__restrict__ bsrc, uint32_t bn, uint32_t log2, const ShEnc e,
                                       uint32_t *fld, uint32_t *rows, uint32_t *pay, uint32_t cap_words, int lane,
                                       uint32_t &bits_out, bool &overflow)
{
    using ST = ShEncStage<ROUNDS>;
    constexpr int QPL = ROUNDS;                                    // quad fields a lane serialises per chunk (ROUNDS * 32 / 32 lanes)
    const int32_t Q = (int32_t)((bn + 3) >> 2);
    const uint32_t kcol = (uint32_t)(Q - 1 - lane) & 31;
    const int32_t mtop = Q - 1 - (int32_t)kcol;
    const uint32_t G = (uint32_t)(Q + 31) >> 5;
    const bool aligned4 = (((uintptr_t)bsrc) & 3) == 0;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t cw = 0, cb = 0, wdone = 0;
    uint32_t *myrow = rows + lane * ST::ROW_STRIDE;
    uint32_t *obuf = fld;
    overflow = false;
    // staging tile: ROUNDS rows x 32 quads (8 bytes each); 16-byte chunk c of row r lives at c ^ key(r, c) (fse_encode128.cuh)
    const uint32_t wchunk = kcol >> 1, whalf = kcol >> 4;
    // reader: lane L takes QPL consecutive quads of the chunk's stream: row = L * QPL / 32, first quad = (L * QPL) % 32
    const uint32_t rrow = ((uint32_t)lane * QPL) >> 5, rq0 = ((uint32_t)lane * QPL) & 31;

    uint32_t sy[ROUNDS];
    const bool first_plain = (bn & 3) == 0;
    auto plain = [&](uint32_t g0) -> bool { return (g0 >= ROUNDS || first_plain) && (uint32_t)Q >= (g0 + ROUNDS) * 32; };
    auto fetch = [&](uint32_t g0) {
        if (!plain(g0)) return;
        if (aligned4) {
            const uint32_t *p32 = reinterpret_cast<const uint32_t *>(bsrc) + (mtop - (int32_t)(g0 << 5));
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) sy[r] = __ldg(p32 - 32 * r);
        } else {
            const uint8_t *p8 = bsrc + 4 * (mtop - (int32_t)(g0 << 5));
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                const uint8_t *q = p8 - 128 * r;
                sy[r] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) | ((uint32_t)__ldg(q + 3) << 24);
            }
        }
    };
    auto store_field = [&](int r, uint2 f) {
        const uint32_t key = ((r & 3) << 1) | whalf;
        *reinterpret_cast<uint2 *>(fld + r * 64 + ((wchunk ^ key) << 2) + ((kcol & 1) << 1)) = f;
    };
    fetch(0);
    for (uint32_t g0 = 0; g0 < G; g0 += ROUNDS) {
        if (plain(g0)) {
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                const uint32_t x = sy[r];
                uint2 f;
                if (r == 0 && g0 == 0) {                           // Encoder::new_first_symbol: no bits
                    s3 = sh_enc_first<NSR>(e, x >> 24);
                    s2 = sh_enc_first<NSR>(e, (x >> 16) & 0xff);
                    s1 = sh_enc_first<NSR>(e, (x >> 8) & 0xff);
                    s0 = sh_enc_first<NSR>(e, x & 0xff);
                    f = make_uint2(0u, 0u);
                } else {
                    uint32_t b3, b2, b1, b0;
                    const uint32_t o3 = s3, o2 = s2, o1 = s1, o0 = s0;
                    s3 = sh_enc_step_at<NSR>(e, sh_tt_addr<3>(e, x), o3, b3);   // decreasing index order: 4m+3 first
                    s2 = sh_enc_step_at<NSR>(e, sh_tt_addr<2>(e, x), o2, b2);
                    s1 = sh_enc_step_at<NSR>(e, sh_tt_addr<1>(e, x), o1, b1);
                    s0 = sh_enc_step_at<NSR>(e, sh_tt_addr<0>(e, x), o0, b0);
                    f = sh_quad_field(o3, b3, o2, b2, o1, b1, o0, b0);
                }
                store_field(r, f);
            }
        } else {
            for (int r = 0; r < ROUNDS; r++) {
                const int32_t m = mtop - (int32_t)((g0 + r) << 5);
                const int32_t i = m < 0 ? -8 : 4 * m;
                uint32_t o3, b3, o2, b2, o1, b1, o0, b0;
                sh_elem_checked<NSR>(bsrc, i + 3, (int32_t)bn, e, s3, o3, b3);
                sh_elem_checked<NSR>(bsrc, i + 2, (int32_t)bn, e, s2, o2, b2);
                sh_elem_checked<NSR>(bsrc, i + 1, (int32_t)bn, e, s1, o1, b1);
                sh_elem_checked<NSR>(bsrc, i, (int32_t)bn, e, s0, o0, b0);
                store_field(r, sh_quad_field(o3, b3, o2, b2, o1, b1, o0, b0));
            }
        }
        if (g0 + ROUNDS < G) fetch(g0 + ROUNDS);
        __syncwarp();
        // pass 2: lane L serialises QPL quad fields that are consecutive in the stream
        BitRowS br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
        {
            const uint32_t rkey_row = (rrow & 3) << 1;
#pragma unroll
            for (int q = 0; q < QPL / 2; q++) {
                const uint32_t c = (rq0 >> 1) + q;                 // 16-byte chunk (two quads) of the row
                const uint32_t key = rkey_row | (c >> 3);
                const uint4 x = *reinterpret_cast<const uint4 *>(fld + rrow * 64 + ((c ^ key) << 2));
                br.put64(x.x, x.y & QUAD_HI_MASK, x.y >> QUAD_LEN_SHIFT);
                br.put64(x.z, x.w & QUAD_HI_MASK, x.w >> QUAD_LEN_SHIFT);
            }
        }
        uint32_t tot = br.finish();
        __syncwarp();
        uint32_t nw = warp_place(myrow, tot, obuf, ST::FLD_WORDS, lane, cw, cb, overflow);
        __syncwarp();
        if (wdone + nw > cap_words) { overflow = true; nw = 0; }
        {
            uint32_t j = lane;
            for (; j + 96 < nw; j += 128) {
                uint32_t w0 = obuf[j], w1 = obuf[j + 32], w2 = obuf[j + 64], w3 = obuf[j + 96];
                pay[wdone + j] = w0; pay[wdone + j + 32] = w1; pay[wdone + j + 64] = w2; pay[wdone + j + 96] = w3;
            }
            for (; j < nw; j += 32) pay[wdone + j] = obuf[j];
        }
        wdone += nw;
        __syncwarp();
    }
    {   // final states 127 .. 0 (fse.rs:248-250), then the marker bit (lib.rs:141,181)
        uint32_t t3 = __shfl_sync(FULL, s3, 31 - lane), t2 = __shfl_sync(FULL, s2, 31 - lane);
        uint32_t t1 = __shfl_sync(FULL, s1, 31 - lane), t0 = __shfl_sync(FULL, s0, 31 - lane);
        const uint32_t mask = (1u << log2) - 1u;
        BitRowS br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
        br.put(t3 & mask, log2);
        br.put(t2 & mask, log2);
        br.put(t1 & mask, log2);
        br.put(t0 & mask, log2);
        if (lane == 31) br.put(1, 1);
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

struct ShEncLayout { uint32_t tab, tt, stage, per_warp, total; };
template <int ROUNDS, int NSR>
__host__ __device__ inline ShEncLayout sh_enc_layout(uint32_t log2, int warps)
{
    ShEncLayout l;
    l.tab = 0;
    l.tt = (1u << log2) * 2 * NSR;
    l.stage = l.tt + 256 * 128;
    l.per_warp = (uint32_t)ShEncStage<ROUNDS>::BYTES;
    l.total = l.stage + l.per_warp * (uint32_t)warps;
    return l;
}

// the CTA copies a table set (reference layout: uint16 table[size], {bits, find_state}[256]) into the bank-replicated form
template <int NSR>
__device__ __forceinline__ void sh_replicate_enc(const uint16_t *__restrict__ tab, const uint2 *__restrict__ tt, uint32_t log2,
                                                 uint8_t *tabR, uint8_t *ttR, int tid, int nthr)
{
    const uint32_t half = 1u << (log2 - 1);
    constexpr uint32_t VPR = NSR / 4;                         // 16-byte vectors per row of NSR copies of one word
    for (uint32_t i = tid; i < half * VPR; i += nthr) {
        const uint32_t j = i / VPR;
        const uint32_t w = (uint32_t)tab[2 * j] | ((uint32_t)tab[2 * j + 1] << 16);
        reinterpret_cast<uint4 *>(tabR)[i] = make_uint4(w, w, w, w);
    }
    for (uint32_t i = tid; i < 256 * 8; i += nthr) {
        const uint32_t w = sh_pack_tt(tt[i >> 3], log2);
        reinterpret_cast<uint4 *>(ttR)[i] = make_uint4(w, w, w, w);
    }
}

// global-table mode: one table for the job; every warp of the CTA codes its own blocks against the CTA's replicated copy
template <int ROUNDS, int NSR>
__global__ void __launch_bounds__(512) k_encode_sh_global(EncArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const uint32_t log2 = a.g.log2;
    const ShEncLayout lay = sh_enc_layout<ROUNDS, NSR>(log2, warps);
    uint8_t *tabR = smem_raw + lay.tab, *ttR = smem_raw + lay.tt;
    uint32_t *fld = reinterpret_cast<uint32_t *>(smem_raw + lay.stage + (size_t)warp * lay.per_warp);
    uint32_t *rows = fld + ShEncStage<ROUNDS>::FLD_WORDS;
    sh_replicate_enc<NSR>(a.g.enc_table, a.g.enc_tt, log2, tabR, ttR, threadIdx.x, blockDim.x);
    const ShEnc e{(uint32_t)__cvta_generic_to_shared(ttR) + 4u * lane,
                  (uint32_t)__cvta_generic_to_shared(tabR) - SH_FS_BIAS * 2u * NSR + 4u * (lane & (NSR - 1))};
    const uint32_t N = 128;
    __shared__ uint32_t cta_next;
    const uint32_t cta_first = (uint32_t)(((unsigned long long)a.nblocks * blockIdx.x) / gridDim.x);
    const uint32_t cta_last = (uint32_t)(((unsigned long long)a.nblocks * (blockIdx.x + 1)) / gridDim.x);
    if (threadIdx.x == 0) cta_next = cta_first;
    __syncthreads();
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(&cta_next, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= cta_last) break;
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        const uint8_t *bsrc = a.src + off;
        uint8_t *bs = a.scratch + (size_t)b * a.stride;
        uint32_t *pay = reinterpret_cast<uint32_t *>(bs + HDR_RESERVE);
        uint32_t hl = 0, pl = 0;
        int st = ST_OK;
        if (bn < N) {                                        // global mode: a short tail is stored raw, no escape
            for (uint32_t i = lane; i < bn; i += 32) bs[i] = bsrc[i];
            hl = bn; st = 1;
        } else {
            uint32_t pbits;
            bool ovf;
            sh_encode_payload_warp<ROUNDS, NSR>(bsrc, bn, log2, e, fld, rows, pay, a.pay_cap_words, lane, pbits, ovf);
            if (ovf) st = ST_CAPACITY;
            else pl = (pbits + 7) >> 3;
        }
        __syncwarp();
        if (lane == 0) { a.status[b] = st; a.hlen[b] = hl; a.plen[b] = pl; }
    }
}

// ------------------------------------------------------------------------------------------
// Segmented per-block mode: one CTA per block.  The block has its own table (one header); its
// segments (segment_size bytes each, independent 128-state streams, header-less like fse.rs:394-421)
// are coded by the warps of the CTA against the CTA's replicated copy of that table.  A builder warp
// normalises and builds the table set of the NEXT block while the others code the current one.
//   stream slot s = block * segs_per_block + k: scratch + s * stride: [0, HDR_RESERVE) the block's header
//   (k == 0 only) or the raw bytes of a short tail segment, [HDR_RESERVE, ...) the payload words.
// ------------------------------------------------------------------------------------------
struct ShBlockMeta { uint32_t log2, hl; int kind; };     // kind 0: coded; 1 / 2: raw / run escape; < 0: error status

constexpr uint32_t SH_BUILD_BYTES = 8192 + 2304;          // tab u16[2048] | tt uint2[256] | norm | cum | spread u8[2048] / ncount rows
struct ShEncBlocksLayout { uint32_t tab, tt, build, meta, stage, per_warp, total; };
template <int ROUNDS, int NSR>
__host__ __device__ inline ShEncBlocksLayout sh_enc_blocks_layout(uint32_t tlmax, int coder_warps)
{
    ShEncBlocksLayout l;
    l.tab = 0;
    l.tt = (1u << tlmax) * 2 * NSR;
    l.build = l.tt + 256 * 128;
    l.meta = l.build + SH_BUILD_BYTES;
    l.stage = l.meta + 32;
    l.per_warp = (uint32_t)ShEncStage<ROUNDS>::BYTES;
    l.total = l.stage + l.per_warp * (uint32_t)coder_warps;
    return l;
}

// builder warp: table set of block b into the build area (reference layout), header into the block's first slot
__device__ __forceinline__ void sh_build_enc_block(const EncArgs &a, uint32_t b, uint8_t *build, ShBlockMeta *meta, int lane)
{
    uint16_t *tab = reinterpret_cast<uint16_t *>(build);
    uint2 *tt = reinterpret_cast<uint2 *>(build + 4096);
    int32_t *norm = reinterpret_cast<int32_t *>(build + 6144);
    uint32_t *cum = reinterpret_cast<uint32_t *>(build + 7168);
    uint8_t *spread = build + 8192;
    uint32_t *rows = reinterpret_cast<uint32_t *>(build + 8192);      // NCount bit strings: dead before the spread is written
    const size_t off = (size_t)b * a.block_size;
    const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
    const uint8_t *bsrc = a.src + off;
    const uint32_t S = a.segs_per_block, s0 = b * S;
    const uint32_t nseg = (bn + a.seg_size - 1) / a.seg_size;
    uint8_t *bs = a.scratch + (size_t)s0 * a.stride;
    uint32_t log2 = 0, table_len = 0, hl = 0;
    int kind = 0;
    int rc = warp_normalize(a.counts + (size_t)b * 256, (uint64_t)bn, a.req_log2, norm, lane, log2, table_len);
    if (rc < 0) {                                            // blocks the reference panics on: escapes (include/fse_b200.h)
        if (table_len <= 1) { if (lane == 0) { bs[0] = 0x0E; bs[1] = 0x00; } hl = 2; kind = 2; }
        else if (bn <= 4) { if ((uint32_t)lane < bn) bs[1 + lane] = bsrc[lane]; if (lane == 0) bs[0] = 0x0F; hl = 1 + bn; kind = 1; }
        else kind = rc;
    } else if (bn < 128) {                                   // fewer symbols than states: stored raw
        for (uint32_t i = lane; i < bn; i += 32) bs[1 + i] = bsrc[i];
        if (lane == 0) bs[0] = 0x0F;
        hl = 1 + bn; kind = 1;
    } else if (log2 > a.tlmax || log2 > SH_TL_MAX) kind = ST_UNSUPPORTED;
#if defined(FSE_DIAG_SKIPBUILD)                               /* timing experiment: tables of the CTA's first block for all its blocks */
    if (b != (uint32_t)(((unsigned long long)a.nblocks * blockIdx.x) / gridDim.x)) { if (lane == 0) { meta->hl = 27; meta->kind = 0; } __syncwarp(); return; }
#endif
    if (rc >= 0 && kind == 0) {
        const uint32_t hbits = warp_ncount_write(norm, log2, table_len, rows, reinterpret_cast<uint32_t *>(bs), lane);
        hl = (hbits + 7) >> 3;
        __syncwarp();
        warp_spread(norm, log2, table_len, spread, cum, tab, lane);
        warp_build_encode(norm, log2, table_len, spread, cum, tab, tt, lane);
    } else {                                                 // the builder owns the index entries of a block that is not coded
        for (uint32_t k = lane; k < nseg; k += 32) {
            a.hlen[s0 + k] = (k == 0 && kind > 0) ? hl : 0u;
            a.plen[s0 + k] = 0;
            a.status[s0 + k] = kind;
        }
    }
    __syncwarp();
    if (lane == 0) { meta->log2 = log2; meta->hl = hl; meta->kind = kind; }
    __syncwarp();
}

template <int ROUNDS, int NSR>
__global__ void __launch_bounds__(544) k_encode_sh_blocks(EncArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int coders = warps - 1;                            // the last warp builds tables
    const ShEncBlocksLayout lay = sh_enc_blocks_layout<ROUNDS, NSR>(a.tlmax, coders);
    uint8_t *tabR = smem_raw + lay.tab, *ttR = smem_raw + lay.tt, *build = smem_raw + lay.build;
    ShBlockMeta *meta = reinterpret_cast<ShBlockMeta *>(smem_raw + lay.meta);
    uint32_t *fld = reinterpret_cast<uint32_t *>(smem_raw + lay.stage + (size_t)(warp < coders ? warp : 0) * lay.per_warp);
    uint32_t *rows = fld + ShEncStage<ROUNDS>::FLD_WORDS;
    const ShEnc e{(uint32_t)__cvta_generic_to_shared(ttR) + 4u * lane,
                  (uint32_t)__cvta_generic_to_shared(tabR) - SH_FS_BIAS * 2u * NSR + 4u * (lane & (NSR - 1))};
    const uint32_t first = (uint32_t)(((unsigned long long)a.nblocks * blockIdx.x) / gridDim.x);
    const uint32_t last = (uint32_t)(((unsigned long long)a.nblocks * (blockIdx.x + 1)) / gridDim.x);
    const bool builder = warp == coders;
    if (builder && first < last) sh_build_enc_block(a, first, build, meta, lane);
    __syncthreads();
    for (uint32_t b = first; b < last; b++) {
        const uint32_t log2 = meta->log2, hl = meta->hl;
        const int kind = meta->kind;
        if (kind == 0)
            sh_replicate_enc<NSR>(reinterpret_cast<const uint16_t *>(build), reinterpret_cast<const uint2 *>(build + 4096), log2,
                                  tabR, ttR, threadIdx.x, blockDim.x);
        __syncthreads();                                     // the replicated tables are complete, the build area is free
        if (builder) {
            if (b + 1 < last) sh_build_enc_block(a, b + 1, build, meta, lane);
        } else if (kind == 0) {
            const size_t off = (size_t)b * a.block_size;
            const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
            const uint32_t nseg = (bn + a.seg_size - 1) / a.seg_size;
            for (uint32_t k = warp; k < nseg; k += coders) {
                const uint32_t slot = b * a.segs_per_block + k;
                const uint32_t so = k * a.seg_size, sn = min(a.seg_size, bn - so);
                const uint8_t *ssrc = a.src + off + so;
                uint8_t *bs = a.scratch + (size_t)slot * a.stride;
                uint32_t h = k == 0 ? hl : 0u, pl = 0;
                int st = ST_OK;
                if (sn < 128) {                              // a short tail segment is stored raw (k > 0: block escapes cover bn < 128)
                    for (uint32_t i = lane; i < sn; i += 32) bs[i] = ssrc[i];
                    h = sn; st = 1;
                } else {
                    uint32_t pbits;
                    bool ovf;
                    sh_encode_payload_warp<ROUNDS, NSR>(ssrc, sn, log2, e, fld, rows, reinterpret_cast<uint32_t *>(bs + HDR_RESERVE),
                                                        a.pay_cap_words, lane, pbits, ovf);
                    if (ovf) { st = ST_CAPACITY; h = 0; }
                    else pl = (pbits + 7) >> 3;
                }
                __syncwarp();
                if (lane == 0) { a.status[slot] = st; a.hlen[slot] = h; a.plen[slot] = pl; }
            }
        }
        __syncthreads();                                     // everyone is done with the tables; the next set is built
    }
}

}  // namespace fsed
