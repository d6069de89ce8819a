// fse_kernels128.cuh -- the 128-state path: one warp per block, four ADJACENT states per lane.
//
// Lane l owns states 4l..4l+3 = the symbol quads (4m+3, 4m+2, 4m+1, 4m) with m == l (mod 32).  Against the
// 64-state path this doubles the independent state chains per lane (the kernels are bound by the latency of
// dependent shared-memory look-ups and by shared-memory wavefronts, not by HBM), halves the per-symbol cost
// of everything that happens once per lane per round (prefix scan, bit fetch, output store, symbol load)
// and lets the decoder use one 32-bit DecodeTransform load per state.  table_log <= 13.
#pragma once
#include "fse_decode64w.cuh"

namespace fsed {

// ------------------------------------------------------------------------------------------
// decode, 128 states.  Shared memory per warp: the Dec64wLayout (4*size + 2 KiB + ring mirror).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 2) k_decode128_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Dec64wLayout lay = dec64w_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint32_t *tab = reinterpret_cast<uint32_t *>(my + lay.tab);
    uint8_t *sym = my + lay.sym;                            // the spread, consumed in place by the table build
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.scratch);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.scratch);          // norm's own array (warp_spread<true>)
    uint32_t *ring = reinterpret_cast<uint32_t *>(my + lay.scratch);   // 256 words + 2 mirror words
    const uint32_t tab_saddr = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t N = 128;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) tab[i] = a.g.dec_table[i];
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
            if (log2 > a.tlmax || log2 > 13) { if (lane == 0) a.status[b] = ST_UNSUPPORTED; continue; }
            warp_spread<true>(norm, log2, table_len, sym, ctr, reinterpret_cast<uint16_t *>(tab), lane);
            warp_build_decode<false>(norm, log2, table_len, sym, ctr, tab, lane);
        }
        if (bn < N) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint8_t *pay = cs + consumed;
        const uint32_t plen = clen - consumed;
        if (plen == 0 || pay[plen - 1] == 0) { if (lane == 0) a.status[b] = ST_NO_MARKER; continue; }
        const uint32_t bias = (uint32_t)((uintptr_t)pay & 3);
        const uint32_t *origin = reinterpret_cast<const uint32_t *>(pay - bias);
        uint32_t cur = (plen - 1) * 8 + ilog2u(pay[plen - 1]) + 8 * bias;
        const uint32_t floor_bits = 8 * bias;
        if (cur - floor_bits < N * log2) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint32_t topq = cur >> 5;
        uint32_t lowq = (topq & ~127u) >= 128 ? (topq & ~127u) - 128 : 0;
        __syncwarp();                                       // the build scratch becomes the ring
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t w = lowq + lane + 32 * k;
            if (w <= topq) {
                uint32_t x = __ldg(origin + w);
                ring[w & 255] = x;
                if ((w & 255) < 2) ring[256 + (w & 255)] = x;   // mirror: ring[256..257] == ring[0..1]
            }
        }
        uint32_t pre[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
        __syncwarp();
        const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
        // up to 52 bits at stream position q (three ring words, 64-bit funnel)
        auto ring_bits64 = [&](uint32_t q) -> uint64_t {
            uint32_t ad = ring_saddr + ((q >> 3) & 0x3fcu);
            uint32_t w0, w1, w2;
            asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                         : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(ad));
            uint32_t s = q & 31;
            return ((uint64_t)__funnelshift_r(w1, w2, s) << 32) | __funnelshift_r(w0, w1, s);
        };
        auto refill = [&]() {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) ring[(lowq - 128 + lane + 32 * k) & 255] = pre[k];
            if (lane < 2 && ((lowq - 128) & 255) == 0) ring[256 + lane] = pre[0];
            lowq -= 128;
#pragma unroll
            for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
            __syncwarp();
        };
        // Decoder::new, fse.rs:349-352: states are read 0, 1, 2, ... from the top of the stack
        uint32_t s0, s1, s2, s3;
        {
            const uint32_t m = (1u << log2) - 1u;
            uint64_t w = ring_bits64(cur - (4 * lane + 4) * log2);
            s3 = (uint32_t)w & m;
            s2 = (uint32_t)(w >> log2) & m;
            s1 = (uint32_t)(w >> (2 * log2)) & m;
            s0 = (uint32_t)(w >> (3 * log2)) & m;
        }
        cur -= N * log2;
        const uint32_t body = bn - N;
        const bool out_aligned = (((uintptr_t)out) & 3) == 0;
        bool bad = false;
        uint32_t i0 = 0;
        // refill when fewer than 56 words lie between the read position and the ring's lowest word (a round takes at most
        // 52); the threshold only changes in refill(), and is 0 (never reached) once the ring holds the first word
        uint32_t thr = lowq ? (lowq + 56) << 5 : 0u;
#define DEC128W_ROUND(STORE)                                                                                          \
        {                                                                                                             \
            if (cur < thr) { refill(); thr = lowq ? (lowq + 56) << 5 : 0u; }                                          \
            uint32_t e0 = lds_u32(tab_saddr + s0 * 4), e1 = lds_u32(tab_saddr + s1 * 4);   /* fse.rs:363-373, four chains */ \
            uint32_t e2 = lds_u32(tab_saddr + s2 * 4), e3 = lds_u32(tab_saddr + s3 * 4);                                \
            uint32_t n0 = e0 >> 24, n1 = e1 >> 24, n2 = e2 >> 24, n3 = e3 >> 24;                                        \
            uint32_t n23 = n2 + n3, n123 = n1 + n23, nbs = n0 + n123;                                                   \
            uint32_t incl = warp_incl_add_pred(nbs);                                                                    \
            uint64_t w = ring_bits64(cur - incl);           /* state 4l's bits are the uppermost of the lane's window */ \
            uint32_t tot = __shfl_sync(FULL, incl, 31);                                                                 \
            if (tot > cur - floor_bits) { bad = true; break; }                                                          \
            s0 = (e0 & 0xffffu) + ((uint32_t)(w >> n123) & ~(0xffffffffu << n0));                                       \
            s1 = (e1 & 0xffffu) + ((uint32_t)(w >> n23) & ~(0xffffffffu << n1));                                        \
            s2 = (e2 & 0xffffu) + ((uint32_t)(w >> n3) & ~(0xffffffffu << n2));                                         \
            s3 = (e3 & 0xffffu) + ((uint32_t)w & ~(0xffffffffu << n3));                                                 \
            uint32_t sy = __byte_perm(__byte_perm(e0, e1, 0x0062), __byte_perm(e2, e3, 0x0062), 0x5410);   /* byte 2 of e0..e3 */ \
            STORE;                                                                                                      \
            cur -= tot;                                                                                                 \
        }
        if (out_aligned) {
            uint32_t *const ow = reinterpret_cast<uint32_t *>(out) + lane;   // the index stays 32 bits wide
            for (; i0 + 128 <= body; i0 += 128) DEC128W_ROUND(ow[i0 >> 2] = sy)
        } else {
            for (; i0 + 128 <= body; i0 += 128)
                DEC128W_ROUND(out[i0 + 4 * lane] = (uint8_t)sy; out[i0 + 4 * lane + 1] = (uint8_t)(sy >> 8);
                              out[i0 + 4 * lane + 2] = (uint8_t)(sy >> 16); out[i0 + 4 * lane + 3] = (uint8_t)(sy >> 24))
        }
#undef DEC128W_ROUND
        if (!bad && i0 < body) {                            // last partial round
            if ((cur >> 5) < lowq + 56 && lowq) refill();
            uint32_t ia = i0 + 4 * lane;
            uint32_t e0 = tab[s0], e1 = tab[s1], e2 = tab[s2], e3 = tab[s3];
            uint32_t n0 = ia < body ? (e0 >> 24) : 0u, n1 = ia + 1 < body ? (e1 >> 24) : 0u;
            uint32_t n2 = ia + 2 < body ? (e2 >> 24) : 0u, n3 = ia + 3 < body ? (e3 >> 24) : 0u;
            uint32_t n23 = n2 + n3, n123 = n1 + n23, nbs = n0 + n123;
            uint32_t incl = warp_incl_add_pred(nbs);
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) bad = true;
            else {
                uint64_t w = ring_bits64(cur - incl);
                if (ia < body) { out[ia] = (uint8_t)(e0 >> 16); s0 = (e0 & 0xffffu) + ((uint32_t)(w >> n123) & ~(0xffffffffu << n0)); }
                if (ia + 1 < body) { out[ia + 1] = (uint8_t)(e1 >> 16); s1 = (e1 & 0xffffu) + ((uint32_t)(w >> n23) & ~(0xffffffffu << n1)); }
                if (ia + 2 < body) { out[ia + 2] = (uint8_t)(e2 >> 16); s2 = (e2 & 0xffffu) + ((uint32_t)(w >> n3) & ~(0xffffffffu << n2)); }
                if (ia + 3 < body) { out[ia + 3] = (uint8_t)(e3 >> 16); s3 = (e3 & 0xffffu) + ((uint32_t)w & ~(0xffffffffu << n3)); }
                cur -= tot;
            }
        }
        if (!bad) {                                         // Decoder::finish, fse.rs:383-385: i in [body, bn), state i % 128
            out[body + ((4 * lane - body) & 127)] = (uint8_t)(tab[s0] >> 16);
            out[body + ((4 * lane + 1 - body) & 127)] = (uint8_t)(tab[s1] >> 16);
            out[body + ((4 * lane + 2 - body) & 127)] = (uint8_t)(tab[s2] >> 16);
            out[body + ((4 * lane + 3 - body) & 127)] = (uint8_t)(tab[s3] >> 16);
        }
        cur -= floor_bits;
        if (bad || cur != 0) st = ST_LENGTH;
        if (lane == 0) a.status[b] = st;
    }
}

}  // namespace fsed
