// fse_encode128.cuh -- 128-state encode: four adjacent states per lane (see fse_kernels128.cuh).
//
// Per chunk of 16 rounds (2 048 symbols): pass 1 runs four independent state chains per lane from one
// 32-bit symbol load per round and stores the quad as ONE field of up to 52 bits (64-bit store, length in the
// top 6 bits); pass 2 transposes: lane L serialises the half row (round L>>1, half L&1) = 16 quad fields that are
// consecutive in the stream into a private bit string; warp_place concatenates the 32 strings (lane order ==
// stream order).  The symbol transforms live in shared memory in one of two forms chosen per block (below).
// Blocks are split evenly over the CTAs; the warps of a CTA take them from a shared counter.
#pragma once
#include "fse_kernels128.cuh"

// FSE_DIAG (development, tools/diag_enc.py): 1 = skip pass 2 / placement / output, 2 = skip the table look-ups of pass 1.
// The output of such a build is garbage; only the kernel time is meaningful.
#ifndef FSE_DIAG
#define FSE_DIAG 0
#endif
namespace fsed {

// Two forms of the per-symbol transform table:
//  * PK = 0 (table_log 12, 13; global-table mode of this kernel): the reference's {bits, find_state} pair (fse.rs:80-84)
//    with find_state pre-scaled to a shared byte address, one 64-bit load per look-up.
//  * PK = 1 (table_log <= 11): the transform re-packed for few ALU-pipe instructions (round 2; the form the CTA-owned
//    tables of fse_shared_enc.cuh introduced):
//      P = H << 12 | (find_state + 2048),   H = (max_bits << 13) - (count << max_bits)
//      t = P + (state << 12)                 one IMAD (FMA pipe)
//      nb = t >> 25                          = (bits + state) >> 16 of fse.rs:228: state - (count << max_bits) lies in
//                                              (-2^12, 2^12), so a 13-bit fraction decides max_bits vs max_bits - 1
//      u = (t & 0xfff) + (state >> nb)       = find_state + 2048 + (state >> nb): the table index, biased by 2048
//    and the emitted bits are never masked out of the state: a funnel shift moves the low nb bits of the raw state into
//    the top of a pair accumulator (quad_field_raw).  One 32-bit load (a whole warp per wavefront instead of half a
////    warp), two copies interleaved by lane parity to halve the lanes per bank.  Against round 1's packed form (p = bits |
//    find_state << 20, which lost to PK = 0 on few-symbol data and was chosen per block): c4 encode 9.48 -> 8.87 ms,
//    c2 0.405 -> 0.389 ms, few-symbol 1 GiB 1.29 -> 1.18 ms (table_log 11), 1.10 -> 1.00 ms (9): used whenever it fits.
constexpr uint32_t TT_REPL_LOG2 = 1;
constexpr uint32_t TT_PACKED_MAX_LOG2 = 11;
constexpr uint32_t SH_FS_BIAS = 2048;
// reference transform {bits, find_state} (fse.rs:165-188) -> P
// A symbol the table does not know (count 0: bits = ((log2 + 1) << 16) - size, fse.rs:170) would code log2 + 1 bits per
// occurrence; it is clamped to log2 so that a quad never exceeds 4 * log2 bits, the size the lane strings are made for
// (such a block is undecodable either way: see fse_b200_set_global_table in include/fse_b200.h).
__device__ __forceinline__ uint32_t sh_pack_tt(uint2 t, uint32_t log2)
{
    uint32_t mbo = (t.x + 65535u) >> 16;                       // bits = (mbo << 16) - y, 0 < y <= 2 * size
    const uint32_t y = (mbo << 16) - t.x;
    mbo = min(mbo, log2);
    const uint32_t H = (mbo << 13) - y;
    return (H << 12) | ((t.y + SH_FS_BIAS) & 0xfffu);
}
// the quad (chains 3, 2, 1, 0 in stream order) as one field: value (hi:lo) right aligned, length in hi[26..31];
// s = the OLD states, b = bits each emits
constexpr uint32_t QUAD_LEN_SHIFT = 26;              // a quad field: 52 value bits, 6 length bits
constexpr uint32_t QUAD_HI_MASK = (1u << QUAD_LEN_SHIFT) - 1u;
__device__ __forceinline__ uint2 quad_field_raw(uint32_t s3, uint32_t b3, uint32_t s2, uint32_t b2, uint32_t s1, uint32_t b1,
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

struct Enc128Tab { uint32_t tt, tab; };    // shared byte addresses: transforms (this lane's copy when packed), next-state
                                           // table (PK = 1: biased by -2 * SH_FS_BIAS)
template <int PK>
__device__ __forceinline__ void enc128_step(const Enc128Tab &e, uint32_t sym, uint32_t &state, uint32_t &v, uint32_t &bo)
{
    enc_step(e.tt, sym, state, v, bo);                         // PK = 0 only; PK = 1 goes through pk_step_at
}
template <int PK>
__device__ __forceinline__ uint32_t enc128_first(const Enc128Tab &e, uint32_t sym)
{
    if (PK) {                                                  // fse.rs:210-218: bo = max_bits, value = count << max_bits
        const uint32_t p = lds_u32(e.tt + (sym << (2 + TT_REPL_LOG2)));
        const uint32_t H = p >> 12;
        const uint32_t mbo = (H + 8191u) >> 13;
        const uint32_t x = ((mbo << 13) - H) >> mbo;
        return lds_u16(e.tab + (((p & 0xfffu) + x) << 1));
    }
    return enc_first64(e.tt, sym);
}
// PK = 1: shared address of the lane's copy of P[byte K of x] (PRMT on the ALU pipe + IMAD on the FMA pipe), and one
// transition (fse.rs:227-239); the caller emits the low nb bits of the OLD state
template <int K>
__device__ __forceinline__ uint32_t pk_tt_addr(const Enc128Tab &e, uint32_t x)
{
    uint32_t a;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(__byte_perm(x, 0u, 0x4440u + K)), "n"(4 << TT_REPL_LOG2), "r"(e.tt));
    return a;
}
__device__ __forceinline__ uint32_t pk_step_at(const Enc128Tab &e, uint32_t tt_addr, uint32_t s, uint32_t &nb)
{
    const uint32_t t = lds_u32(tt_addr) + (s << 12);
    nb = t >> 25;
    const uint32_t u = (t & 0xfffu) + (s >> nb);
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a) : "r"(u), "r"(e.tab));
    return lds_u16(a);
}
__device__ __forceinline__ void pk_elem_checked(const uint8_t *__restrict__ bsrc, int32_t i, int32_t bn, const Enc128Tab &e,
                                                uint32_t &s, uint32_t &sold, uint32_t &nb)
{
    sold = 0; nb = 0;
    if (i < 0 || i >= bn) return;
    const uint32_t sym = __ldg(bsrc + i);
    if (i >= bn - 128) s = enc128_first<1>(e, sym);
    else { sold = s; s = pk_step_at(e, e.tt + (sym << (2 + TT_REPL_LOG2)), s, nb); }
}

// element classes of a symbol index for N = 128 (cf. enc_element_checked)
template <int PK>
__device__ __forceinline__ void enc_element_checked128(const uint8_t *__restrict__ bsrc, int32_t i, int32_t bn, const Enc128Tab &tt_saddr,
                                                       uint32_t &state, uint32_t &v, uint32_t &bo)
{
    v = 0; bo = 0;
    if (i < 0 || i >= bn) return;
    uint32_t sym = __ldg(bsrc + i);
    if (i >= bn - 128) state = enc128_first<PK>(tt_saddr, sym);
    else enc128_step<PK>(tt_saddr, sym, state, v, bo);
}

// BitRow with the word emission written as predicated PTX: 9 instructions per field instead of the 13 the
// compiler makes of the C++ version (it materialises every conditional update as add + move).
struct BitRowS {
    uint32_t base, wp, lo, pos;      // shared byte addresses of the row start / next word, accumulator, bits held
    __device__ __forceinline__ void init(uint32_t *row, uint32_t carry_word, uint32_t carry_bits)
    {
        base = wp = (uint32_t)__cvta_generic_to_shared(row);
        lo = carry_word; pos = carry_bits;
    }
    // v < 2^nb, nb <= 26, pos < 32
    __device__ __forceinline__ void put(uint32_t v, uint32_t nb)
    {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t, h;\n\t"
                     "shl.b32 t, %3, %1;\n\t"
                     "or.b32 %0, %0, t;\n\t"
                     "shf.l.clamp.b32 h, %3, 0, %1;\n\t"      // bits of v that spill past 32
                     "add.u32 %1, %1, %4;\n\t"
                     "setp.ge.u32 p, %1, 32;\n\t"
                     "@p st.shared.u32 [%2], %0;\n\t"
                     "@p add.u32 %2, %2, 4;\n\t"
                     "@p mov.u32 %0, h;\n\t"
                     "and.b32 %1, %1, 31;\n\t}"
                     : "+r"(lo), "+r"(pos), "+r"(wp) : "r"(v), "r"(nb) : "memory");
    }
    // a quad field: value (fhi:flo) < 2^nb, nb <= 52, pos < 32: 0, 1 or 2 words are completed per call
    __device__ __forceinline__ void put64(uint32_t flo, uint32_t fhi, uint32_t nb)
    {
        asm volatile("{\n\t.reg .pred p1, p2;\n\t.reg .u32 t, w1, w2, np, c;\n\t"
                     "shl.b32 t, %3, %1;\n\t"
                     "or.b32 %0, %0, t;\n\t"
                     "shf.l.clamp.b32 w1, %3, %4, %1;\n\t"
                     "shf.l.clamp.b32 w2, %4, 0, %1;\n\t"
                     "add.u32 np, %1, %5;\n\t"
                     "shr.u32 c, np, 5;\n\t"
                     "setp.ne.u32 p1, c, 0;\n\t"
                     "setp.eq.u32 p2, c, 2;\n\t"
                     "@p1 st.shared.u32 [%2], %0;\n\t"
                     "@p2 st.shared.u32 [%2+4], w1;\n\t"
                     "mad.lo.u32 %2, c, 4, %2;\n\t"
                     "@p1 mov.u32 %0, w1;\n\t"
                     "@p2 mov.u32 %0, w2;\n\t"
                     "and.b32 %1, np, 31;\n\t}"
                     : "+r"(lo), "+r"(pos), "+r"(wp) : "r"(flo), "r"(fhi), "r"(nb) : "memory");
    }
    __device__ __forceinline__ uint32_t finish()
    {
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(wp), "r"(lo) : "memory");
        return ((wp - base) << 3) + pos;
    }
};

template <int PK>
__device__ void encode128_payload_warp(const uint8_t *__restrict__ bsrc, uint32_t bn, uint32_t log2, const Enc128Tab tt_saddr,
                                       uint32_t *fld, uint32_t *rows, uint32_t *pay, uint32_t cap_words, int lane,
                                       uint32_t &bits_out, bool &overflow)
{
    const int32_t Q = (int32_t)((bn + 3) >> 2);                    // quads (4m+3 .. 4m)
    const uint32_t kcol = (uint32_t)(Q - 1 - lane) & 31;           // my quad's stream position in a round
    const int32_t mtop = Q - 1 - (int32_t)kcol;                    // my quad in round 0
    const uint32_t G = (uint32_t)(Q + 31) >> 5;                    // rounds of 32 quads
    const bool aligned4 = (((uintptr_t)bsrc) & 3) == 0;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;                       // states 4*lane .. 4*lane+3
    uint32_t cw = 0, cb = 0, wdone = 0;
    uint32_t *myrow = rows + lane * ROW_STRIDE64;
    uint32_t *obuf = fld;
    overflow = false;
    // staging tile: 16 rows x 64 pair fields; 16-byte chunk c of row r lives at c ^ (((r & 3) << 1) | (c >> 3)) (low 3 bits)
    const uint32_t wchunk = kcol >> 1, whalf = kcol >> 4;
    const uint32_t rrow = (uint32_t)lane >> 1, rhalf = (uint32_t)lane & 1, rkey = ((rrow & 3) << 1) | rhalf;

    uint32_t sy[16];
    // a chunk whose 16 rounds are all inside the block runs without per-element checks; chunk 0 qualifies when
    // the block length is a multiple of 4, so that round 0 is exactly the 128 state-initialising symbols
    const bool first_plain = (bn & 3) == 0;
    auto plain = [&](uint32_t g0) -> bool { return (g0 >= 16 || first_plain) && (uint32_t)Q >= (g0 + 16) * 32; };
    // only plain chunks read sy[] (the checked path loads its own bytes): one 32-bit load per quad when the
    // block is 4-byte aligned, four byte loads otherwise
    auto fetch = [&](uint32_t g0) {
        if (!plain(g0)) return;
        if (aligned4) {
            const uint32_t *p32 = reinterpret_cast<const uint32_t *>(bsrc) + (mtop - (int32_t)(g0 << 5));
#pragma unroll
            for (int r = 0; r < 16; r++) sy[r] = __ldg(p32 - 32 * r);
        } else {
            const uint8_t *p8 = bsrc + 4 * (mtop - (int32_t)(g0 << 5));
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const uint8_t *q = p8 - 128 * r;
                sy[r] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) | ((uint32_t)__ldg(q + 3) << 24);
            }
        }
    };
    fetch(0);
    for (uint32_t g0 = 0; g0 < G; g0 += 16) {
        if (plain(g0)) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
                uint32_t v3, b3, v2, b2, v1, b1, v0, b0;
                const uint32_t x = sy[r];
                uint2 f;
                if (r == 0 && g0 == 0) {                           // Encoder::new_first_symbol: no bits
                    s3 = enc128_first<PK>(tt_saddr, x >> 24);
                    s2 = enc128_first<PK>(tt_saddr, (x >> 16) & 0xff);
                    s1 = enc128_first<PK>(tt_saddr, (x >> 8) & 0xff);
                    s0 = enc128_first<PK>(tt_saddr, x & 0xff);
                    f = make_uint2(0u, 0u);
                } else if (PK) {
                    const uint32_t o3 = s3, o2 = s2, o1 = s1, o0 = s0;
                    s3 = pk_step_at(tt_saddr, pk_tt_addr<3>(tt_saddr, x), o3, b3);   // decreasing index order: 4m+3 first
                    s2 = pk_step_at(tt_saddr, pk_tt_addr<2>(tt_saddr, x), o2, b2);
                    s1 = pk_step_at(tt_saddr, pk_tt_addr<1>(tt_saddr, x), o1, b1);
                    s0 = pk_step_at(tt_saddr, pk_tt_addr<0>(tt_saddr, x), o0, b0);
                    f = quad_field_raw(o3, b3, o2, b2, o1, b1, o0, b0);
                } else {
#if FSE_DIAG == 2                                               // timing experiment: no table look-ups
                    v3 = x >> 27; b3 = 5; v2 = (x >> 16) & 31; b2 = 5; v1 = (x >> 8) & 31; b1 = 5; v0 = x & 31; b0 = 6;
#else
                    enc128_step<PK>(tt_saddr, x >> 24, s3, v3, b3);       // decreasing index order: 4m+3 first
                    enc128_step<PK>(tt_saddr, (x >> 16) & 0xff, s2, v2, b2);
                    enc128_step<PK>(tt_saddr, (x >> 8) & 0xff, s1, v1, b1);
                    enc128_step<PK>(tt_saddr, x & 0xff, s0, v0, b0);
#endif
                    // the quad as one field of <= 52 bits: (hi:lo), length in hi[26..31]
                    const uint32_t p1 = v3 | (v2 << b3), n1 = b3 + b2, p0 = v1 | (v0 << b1);
                    f.x = p1 | (p0 << n1);
                    f.y = __funnelshift_l(p0, 0u, n1) | ((n1 + b1 + b0) << QUAD_LEN_SHIFT);
                }
                const uint32_t key = ((r & 3) << 1) | whalf;
                *reinterpret_cast<uint2 *>(fld + r * 64 + ((wchunk ^ key) << 2) + ((kcol & 1) << 1)) = f;
            }
        } else {
            for (int r = 0; r < 16; r++) {
                int32_t m = mtop - (int32_t)((g0 + r) << 5);
                int32_t i = m < 0 ? -8 : 4 * m;
                uint32_t v3, b3, v2, b2, v1, b1, v0, b0;
                uint2 f;
                if (PK) {
                    pk_elem_checked(bsrc, i + 3, (int32_t)bn, tt_saddr, s3, v3, b3);     // v = the old state here
                    pk_elem_checked(bsrc, i + 2, (int32_t)bn, tt_saddr, s2, v2, b2);
                    pk_elem_checked(bsrc, i + 1, (int32_t)bn, tt_saddr, s1, v1, b1);
                    pk_elem_checked(bsrc, i, (int32_t)bn, tt_saddr, s0, v0, b0);
                    f = quad_field_raw(v3, b3, v2, b2, v1, b1, v0, b0);
                } else {
                    enc_element_checked128<PK>(bsrc, i + 3, (int32_t)bn, tt_saddr, s3, v3, b3);
                    enc_element_checked128<PK>(bsrc, i + 2, (int32_t)bn, tt_saddr, s2, v2, b2);
                    enc_element_checked128<PK>(bsrc, i + 1, (int32_t)bn, tt_saddr, s1, v1, b1);
                    enc_element_checked128<PK>(bsrc, i, (int32_t)bn, tt_saddr, s0, v0, b0);
                    // the quad as one field of <= 52 bits: (hi:lo), length in hi[26..31]
                    const uint32_t p1 = v3 | (v2 << b3), n1 = b3 + b2, p0 = v1 | (v0 << b1);
                    f.x = p1 | (p0 << n1);
                    f.y = __funnelshift_l(p0, 0u, n1) | ((n1 + b1 + b0) << QUAD_LEN_SHIFT);
                }
                const uint32_t key = ((r & 3) << 1) | whalf;
                *reinterpret_cast<uint2 *>(fld + r * 64 + ((wchunk ^ key) << 2) + ((kcol & 1) << 1)) = f;
            }
        }
        if (g0 + 16 < G) fetch(g0 + 16);
        __syncwarp();
#if FSE_DIAG == 1                                               // timing experiment: no transpose / placement / output
        wdone += 300; continue;
#endif
        // pass 2: lane L serialises half a round: 32 merged pairs that are consecutive in the stream
        BitRowS br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
#pragma unroll 2
        for (int q = 0; q < 8; q++) {
            uint4 x = *reinterpret_cast<const uint4 *>(fld + rrow * 64 + (((rhalf << 3) | (q ^ rkey)) << 2));
            br.put64(x.x, x.y & QUAD_HI_MASK, x.y >> QUAD_LEN_SHIFT);
            br.put64(x.z, x.w & QUAD_HI_MASK, x.w >> QUAD_LEN_SHIFT);
        }
        uint32_t tot = br.finish();
        __syncwarp();
        uint32_t nw = warp_place(myrow, tot, obuf, 1024, lane, cw, cb, overflow);
        __syncwarp();
        if (wdone + nw > cap_words) { overflow = true; nw = 0; }
        {
            uint32_t j = lane;
            for (; j + 96 < nw; j += 128) {                    // whole 128-byte lines, four in flight
                uint32_t w0 = obuf[j], w1 = obuf[j + 32], w2 = obuf[j + 64], w3 = obuf[j + 96];
                pay[wdone + j] = w0; pay[wdone + j + 32] = w1; pay[wdone + j + 64] = w2; pay[wdone + j + 96] = w3;
            }
            for (; j < nw; j += 32) pay[wdone + j] = obuf[j];
        }
        wdone += nw;
        __syncwarp();
    }
    // final states 127 .. 0 (fse.rs:248-250), then the marker bit (lib.rs:141,181):
    // stream position L holds states 127-4L .. 124-4L, all owned by lane 31-L
    {
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

__global__ void __launch_bounds__(512) k_encode128_blocks(EncArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Enc64Layout lay = enc64_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint16_t *tab = reinterpret_cast<uint16_t *>(my + lay.tab);
    uint2 *tt = reinterpret_cast<uint2 *>(my + lay.tt);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(my + lay.work);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.work + 1024);
    uint32_t *cum = reinterpret_cast<uint32_t *>(my + lay.work + 2048);
    uint8_t *spread = my + lay.work + 3072;
    uint32_t *fld = reinterpret_cast<uint32_t *>(my + lay.work);
    uint32_t *rows = reinterpret_cast<uint32_t *>(my + lay.rows);
    const uint32_t tab_saddr = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t tt_saddr = (uint32_t)__cvta_generic_to_shared(tt);
    const uint32_t N = 128;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) tab[i] = a.g.enc_table[i];
        for (uint32_t i = lane; i < 256; i += 32) {          // global mode keeps the 64-bit form
            uint2 t = a.g.enc_tt[i];
            t.y = tab_saddr + 2u * t.y;
            tt[i] = t;
        }
        __syncwarp();
    }

    // blocks are split evenly over the CTAs (every SM gets the same share whatever the warp count is); inside a CTA
    // the warps take the next block from a shared counter
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
        uint32_t *hdr_words = reinterpret_cast<uint32_t *>(bs);
        uint32_t *pay = reinterpret_cast<uint32_t *>(bs + HDR_RESERVE);
        uint32_t log2 = glog2;
        uint32_t hl = 0, pl = 0;                             // header / payload bytes of this block (warp uniform)
        bool packed = false;                                 // form of this block's symbol transforms
        int st = ST_OK;
        __syncwarp();
        do {
            if (!a.global_mode) {
#pragma unroll
                for (int k = 0; k < 8; k++) cnt[k * 32 + lane] = a.counts[(size_t)b * 256 + k * 32 + lane];
                __syncwarp();
                uint32_t table_len;
                int rc = warp_normalize(cnt, (uint64_t)bn, a.req_log2, norm, lane, log2, table_len);
                if (rc < 0) {
                    // blocks the reference panics on: stored with an escape byte (include/fse_b200.h)
                    if (table_len <= 1) { if (lane == 0) { bs[0] = 0x0E; bs[1] = 0x00; } hl = 2; st = 2; }
                    else if (bn <= 4) { if ((uint32_t)lane < bn) bs[1 + lane] = bsrc[lane]; if (lane == 0) bs[0] = 0x0F; hl = 1 + bn; st = 1; }
                    else st = rc;
                    break;
                }
                if (bn < N) {                                // fewer symbols than states: stored raw
                    for (uint32_t i = lane; i < bn; i += 32) bs[1 + i] = bsrc[i];
                    if (lane == 0) bs[0] = 0x0F;
                    hl = 1 + bn; st = 1;
                    break;
                }
                if (log2 > a.tlmax || log2 > 13) { st = ST_UNSUPPORTED; break; }
                uint32_t hbits = warp_ncount_write(norm, log2, table_len, rows, hdr_words, lane);
                hl = (hbits + 7) >> 3;
                warp_spread(norm, log2, table_len, spread, cum, tab, lane);
                warp_build_encode(norm, log2, table_len, spread, cum, tab, tt, lane);
                packed = log2 <= TT_PACKED_MAX_LOG2;          // the re-packed form wins on every distribution measured (below)
                if (packed) {                                // two interleaved copies of the 32-bit form, in place
                    uint32_t pk[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) pk[k] = sh_pack_tt(tt[k * 32 + lane], log2);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 8; k++)
#pragma unroll
                        for (int r = 0; r < (1 << TT_REPL_LOG2); r++)
                            reinterpret_cast<uint32_t *>(tt)[((k * 32 + lane) << TT_REPL_LOG2) + r] = pk[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 8; k++) {            // pre-scale find_state to a shared-memory byte address
                        uint2 t = tt[k * 32 + lane];
                        t.y = tab_saddr + 2u * t.y;
                        tt[k * 32 + lane] = t;
                    }
                }
                __syncwarp();
            } else if (bn < N) {                             // global mode: a short tail is stored raw, no escape
                for (uint32_t i = lane; i < bn; i += 32) bs[i] = bsrc[i];
                hl = bn; st = 1;
                break;
            }
            uint32_t pbits;
            bool ovf;
                        if (packed) {
                const Enc128Tab et{tt_saddr + 4u * ((uint32_t)lane & ((1u << TT_REPL_LOG2) - 1u)), tab_saddr - 2u * SH_FS_BIAS};
                encode128_payload_warp<1>(bsrc, bn, log2, et, fld, rows, pay, a.pay_cap_words, lane, pbits, ovf);
            } else {
                const Enc128Tab et{tt_saddr, tab_saddr};
                encode128_payload_warp<0>(bsrc, bn, log2, et, fld, rows, pay, a.pay_cap_words, lane, pbits, ovf);
            }
            if (ovf) { st = ST_CAPACITY; hl = 0; pl = 0; }
            else {
                pl = (pbits + 7) >> 3;
                if (!a.global_mode && warp_raw_if_expands(a.flags, hl, pl, bs, bsrc, bn, lane)) { hl = 1; pl = bn; st = 1; }
            }
        } while (0);
        __syncwarp();
        if (lane == 0) a.status[b] = st;
        if (lane == 0) { a.hlen[b] = hl; a.plen[b] = pl; }
    }
}

}  // namespace fsed
