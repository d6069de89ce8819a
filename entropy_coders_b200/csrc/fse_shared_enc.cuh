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
constexpr uint32_t SH_FS_BIAS = 2048;

// reference transform {bits, find_state} (fse.rs:165-188) -> P
__device__ __forceinline__ uint32_t sh_pack_tt(uint2 t)
{
    const uint32_t mbo = (t.x + 65535u) >> 16;                 // bits = (mbo << 16) - y, 0 < y <= 2 * size
    const uint32_t y = (mbo << 16) - t.x;
    const uint32_t H = (mbo << 13) - y;
    return (H << 12) | ((t.y + SH_FS_BIAS) & 0xfffu);
}

struct ShEnc { uint32_t ttl, tabl; };     // ttR + lane * 4;  tabR - SH_FS_BIAS * 64 + lane * 4

__device__ __forceinline__ uint32_t sh_tab_addr(const ShEnc &e, uint32_t u)
{
    return e.tabl + ((u >> 1) << 7) + ((u & 1u) << 1);
}
// one transition (fse.rs:227-239): the caller emits the low nb bits of the OLD state
__device__ __forceinline__ uint32_t sh_enc_step(const ShEnc &e, uint32_t sym, uint32_t s, uint32_t &nb)
{
    const uint32_t t = lds_u32(e.ttl + (sym << 7)) + (s << 12);
    nb = t >> 25;
    const uint32_t u = (t & 0xfffu) + (s >> nb);
    return lds_u16(sh_tab_addr(e, u));
}
// Encoder::new_first_symbol (fse.rs:210-218): bo = max_bits, value = count << max_bits, so the state is table[total]
__device__ __forceinline__ uint32_t sh_enc_first(const ShEnc &e, uint32_t sym)
{
    const uint32_t p = lds_u32(e.ttl + (sym << 7));
    const uint32_t H = p >> 12;
    const uint32_t mbo = (H + 8191u) >> 13;
    const uint32_t x = ((mbo << 13) - H) >> mbo;
    return lds_u16(sh_tab_addr(e, (p & 0xfffu) + x));
}

// the quad (chains 3, 2, 1, 0 in stream order) as one field: value (hi:lo) right aligned, length in hi[26..31]
__device__ __forceinline__ uint2 sh_quad_field(uint32_t s3, uint32_t b3, uint32_t s2, uint32_t b2, uint32_t s1, uint32_t b1,
                                               uint32_t s0, uint32_t b0)
{
    // a pair accumulates top aligned: funnel the low b bits of the raw state in from above, no masks
    uint32_t pa = __funnelshift_r(0u, s3, b3);
    pa = __funnelshift_r(pa, s2, b2);
    uint32_t pb = __funnelshift_r(0u, s1, b1);
    pb = __funnelshift_r(pb, s0, b0);
    const uint32_t na = b3 + b2, nbb = b1 + b0;
    const uint32_t va = __funnelshift_r(pa, 0u, 0u - na);      // pa >> (32 - na); na == 0: pa == 0
    const uint32_t vb = __funnelshift_r(pb, 0u, 0u - nbb);
    uint2 f;
    f.x = va | (vb << na);
    f.y = __funnelshift_l(vb, 0u, na) | ((na + nbb) << QUAD_LEN_SHIFT);
    return f;
}

template <int ROUNDS> struct ShEncStage {
    static constexpr int FLD_WORDS = ROUNDS * 64;                             // ROUNDS rows of 32 quad fields
    static constexpr int ROW_STRIDE = (ROUNDS == 16) ? 29 : 15;               // words per lane string, odd
    static constexpr int ROWS_WORDS = 32 * ROW_STRIDE;
    static constexpr int BYTES = (FLD_WORDS + ROWS_WORDS) * 4;
};

__device__ __forceinline__ void sh_elem_checked(const uint8_t *__restrict__ bsrc, int32_t i, int32_t bn, const ShEnc &e,
                                                uint32_t &s, uint32_t &sold, uint32_t &nb)
{
    sold = 0; nb = 0;
    if (i < 0 || i >= bn) return;
    const uint32_t sym = __ldg(bsrc + i);
    if (i >= bn - 128) s = sh_enc_first(e, sym);
    else { sold = s; s = sh_enc_step(e, sym, s, nb); }
}

// 128-state payload of one stream (block or segment) by one warp; the structure of encode128_payload_warp with the
// shared tables.  ROUNDS = rounds per chunk (16: lane L serialises half a round; 8: a quarter of a round).
template <int ROUNDS>
__device__ void sh_encode_payload_warp(const uint8_t *__restrict__ bsrc, uint32_t bn, uint32_t log2, const ShEnc e,
                                       uint32_t *fld, uint32_t *rows, uint32_t *pay, uint32_t cap_words, int lane,
                                       uint32_t &bits_out, bool &overflow)
{
    using ST = ShEncStage<ROUNDS>;
    constexpr int QPL = ROUNDS / 2;                                // quad fields a lane serialises per chunk
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
                    s3 = sh_enc_first(e, x >> 24);
                    s2 = sh_enc_first(e, (x >> 16) & 0xff);
                    s1 = sh_enc_first(e, (x >> 8) & 0xff);
                    s0 = sh_enc_first(e, x & 0xff);
                    f = make_uint2(0u, 0u);
                } else {
                    uint32_t b3, b2, b1, b0;
                    const uint32_t o3 = s3, o2 = s2, o1 = s1, o0 = s0;
                    s3 = sh_enc_step(e, x >> 24, o3, b3);          // decreasing index order: 4m+3 first
                    s2 = sh_enc_step(e, (x >> 16) & 0xff, o2, b2);
                    s1 = sh_enc_step(e, (x >> 8) & 0xff, o1, b1);
                    s0 = sh_enc_step(e, x & 0xff, o0, b0);
                    f = sh_quad_field(o3, b3, o2, b2, o1, b1, o0, b0);
                }
                store_field(r, f);
            }
        } else {
            for (int r = 0; r < ROUNDS; r++) {
                const int32_t m = mtop - (int32_t)((g0 + r) << 5);
                const int32_t i = m < 0 ? -8 : 4 * m;
                uint32_t o3, b3, o2, b2, o1, b1, o0, b0;
                sh_elem_checked(bsrc, i + 3, (int32_t)bn, e, s3, o3, b3);
                sh_elem_checked(bsrc, i + 2, (int32_t)bn, e, s2, o2, b2);
                sh_elem_checked(bsrc, i + 1, (int32_t)bn, e, s1, o1, b1);
                sh_elem_checked(bsrc, i, (int32_t)bn, e, s0, o0, b0);
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
template <int ROUNDS>
__host__ __device__ inline ShEncLayout sh_enc_layout(uint32_t log2, int warps)
{
    ShEncLayout l;
    l.tab = 0;
    l.tt = (1u << log2) * 64;
    l.stage = l.tt + 256 * 128;
    l.per_warp = (uint32_t)ShEncStage<ROUNDS>::BYTES;
    l.total = l.stage + l.per_warp * (uint32_t)warps;
    return l;
}

// the CTA copies a table set (reference layout: uint16 table[size], {bits, find_state}[256]) into the bank-replicated form
__device__ __forceinline__ void sh_replicate_enc(const uint16_t *__restrict__ tab, const uint2 *__restrict__ tt, uint32_t log2,
                                                 uint8_t *tabR, uint8_t *ttR, int tid, int nthr)
{
    const uint32_t half = 1u << (log2 - 1);
    // a row = 128 bytes = eight 16-byte vectors holding the same word
    for (uint32_t i = tid; i < half * 8; i += nthr) {
        const uint32_t j = i >> 3;
        const uint32_t w = (uint32_t)tab[2 * j] | ((uint32_t)tab[2 * j + 1] << 16);
        reinterpret_cast<uint4 *>(tabR)[i] = make_uint4(w, w, w, w);
    }
    for (uint32_t i = tid; i < 256 * 8; i += nthr) {
        const uint32_t w = sh_pack_tt(tt[i >> 3]);
        reinterpret_cast<uint4 *>(ttR)[i] = make_uint4(w, w, w, w);
    }
}

// global-table mode: one table for the job; every warp of the CTA codes its own blocks against the CTA's replicated copy
template <int ROUNDS>
__global__ void __launch_bounds__(512) k_encode_sh_global(EncArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const uint32_t log2 = a.g.log2;
    const ShEncLayout lay = sh_enc_layout<ROUNDS>(log2, warps);
    uint8_t *tabR = smem_raw + lay.tab, *ttR = smem_raw + lay.tt;
    uint32_t *fld = reinterpret_cast<uint32_t *>(smem_raw + lay.stage + (size_t)warp * lay.per_warp);
    uint32_t *rows = fld + ShEncStage<ROUNDS>::FLD_WORDS;
    sh_replicate_enc(a.g.enc_table, a.g.enc_tt, log2, tabR, ttR, threadIdx.x, blockDim.x);
    const ShEnc e{(uint32_t)__cvta_generic_to_shared(ttR) + 4u * lane,
                  (uint32_t)__cvta_generic_to_shared(tabR) - SH_FS_BIAS * 64u + 4u * lane};
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
            sh_encode_payload_warp<ROUNDS>(bsrc, bn, log2, e, fld, rows, pay, a.pay_cap_words, lane, pbits, ovf);
            if (ovf) st = ST_CAPACITY;
            else pl = (pbits + 7) >> 3;
        }
        __syncwarp();
        if (lane == 0) { a.status[b] = st; a.hlen[b] = hl; a.plen[b] = pl; }
    }
}

}  // namespace fsed
