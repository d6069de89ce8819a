// fse_decode128c.cuh -- 128-state decode over the compact tables of fse_decode64c.cuh (u16 entry + symbol
// array, 8 KiB per warp at table_log 11 = 28 blocks in flight per SM); table_log <= 12.
#pragma once
#include "fse_kernels128.cuh"

namespace fsed {

#ifndef FSE_DEC_TMA
#define FSE_DEC_TMA 1   /* stage the payload ring with cp.async.bulk (TMA) + mbarrier instead of register-prefetched loads */
#endif

// ---- 1-D bulk async copy (TMA) global -> shared with mbarrier completion (SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // earlier generic reads of the slot are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// bounded: a copy that never lands must not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

__global__ void __launch_bounds__(512, 2) k_decode128c_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Dec64cLayout lay = dec64c_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint16_t *tab = reinterpret_cast<uint16_t *>(my + lay.tab);
    uint8_t *sym = my + lay.sym;                            // the spread = the symbol of every cell
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.scratch);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.scratch);          // norm's own array (warp_spread<true>)
    uint32_t *ring = reinterpret_cast<uint32_t *>(my + lay.scratch);   // 256 words + 2 mirror words
    const uint32_t tab_saddr = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t sym_saddr = (uint32_t)__cvta_generic_to_shared(sym);
#if FSE_DEC_TMA
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(my + lay.total - 16);   // one mbarrier per warp, lives across blocks
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    uint32_t par = 0;
#endif
    const uint32_t N = 128;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) {
            uint32_t e = a.g.dec_table[i];
            tab[i] = (uint16_t)((e & 0xfffu) | ((e >> 24) << 12));
            sym[i] = (uint8_t)(e >> 16);
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
            const int rc = warp_ncount_read(cs, clen, reinterpret_cast<uint32_t *>(tab), norm, lane, log2, table_len, consumed);
            if (rc < 0) { if (lane == 0) a.status[b] = rc; continue; }
            if (log2 > a.tlmax || log2 > 12) { if (lane == 0) a.status[b] = ST_UNSUPPORTED; continue; }
            warp_spread<true>(norm, log2, table_len, sym, ctr, tab, lane);
            warp_build_decode16<false>(norm, log2, table_len, sym, ctr, tab, lane);
        }
        if (bn < N) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint8_t *pay = cs + consumed;
        const uint32_t plen = clen - consumed;
        if (plen == 0 || pay[plen - 1] == 0) { if (lane == 0) a.status[b] = ST_NO_MARKER; continue; }
#if FSE_DEC_TMA
        const uint32_t bias = (uint32_t)((uintptr_t)pay & 15);     // bulk copies need 16-byte aligned global chunks
#else
        const uint32_t bias = (uint32_t)((uintptr_t)pay & 3);
#endif
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
#if !FSE_DEC_TMA
        uint32_t pre[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
#endif
        __syncwarp();
        const uint32_t ring_saddr = (uint32_t)__cvta_generic_to_shared(ring);
        // up to 52 bits at stream position q (three ring words, 64-bit funnel)
        auto ring_lohi = [&](uint32_t q, uint32_t &lo, uint32_t &hi) {
            uint32_t ad = ring_saddr + ((q >> 3) & 0x3fcu);
            uint32_t w0, w1, w2;
            asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                         : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(ad));
            uint32_t s = q & 31;
            lo = __funnelshift_r(w0, w1, s);
            hi = __funnelshift_r(w1, w2, s);
        };
        auto ring_bits64 = [&](uint32_t q) -> uint64_t {
            uint32_t lo, hi;
            ring_lohi(q, lo, hi);
            return ((uint64_t)hi << 32) | lo;
        };
#if FSE_DEC_TMA
        // The ring holds words [lowq, lowq+256).  As soon as the upper half is dead, one lane starts a 512-byte
        // bulk copy of the next lower 128 words into it; the warp waits on the mbarrier only when it gets there.
        uint32_t pending = 0, tma_ok = 1;                   // 32-bit flags, touched on the rare paths only
        auto stage = [&]() {
            if ((cur >> 5) + 3 >= lowq + 128 || !lowq) return;     // one test per round: nothing to do in the upper half
            if (!pending) {
                __syncwarp();
                if (lane == 0) bulk_g2s(ring_saddr + (((lowq - 128) & 255) << 2), origin + (lowq - 128), 512, bar);
                pending = 1;
            }
            if ((cur >> 5) < lowq + 56) {                          // a round takes at most 52 words
                if (!mbar_wait(bar, par)) tma_ok = 0;
                par ^= 1;
                lowq -= 128;
                pending = 0;
                if ((lowq & 255) == 0) {                            // slots 0 and 1 changed: refresh their mirror
                    __syncwarp();
                    if (lane < 2) ring[256 + lane] = ring[lane];
                }
                __syncwarp();
            }
        };
#else
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
        auto stage = [&]() { if ((cur >> 5) < lowq + 56 && lowq) refill(); };   // a round takes at most 52 words
#endif
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
        uint32_t bad = 0;
        uint32_t i0 = 0;
        // one full round: fse.rs:363-373 on four chains per lane.  new_state + bits is an OR: the entry's base is a multiple of
        // 1 << num_bits.  The four fields come from three funnel shifts of the lane's 64-bit window (n23 <= 24).
#define DEC128C_ROUND(STORE)                                                                                                  \
        {                                                                                                                     \
            stage();                                                                                                          \
            const uint32_t e0 = lds_u16(tab_saddr + s0 * 2), e1 = lds_u16(tab_saddr + s1 * 2);                                  \
            const uint32_t e2 = lds_u16(tab_saddr + s2 * 2), e3 = lds_u16(tab_saddr + s3 * 2);                                  \
            const uint32_t y0 = lds_u8(sym_saddr + s0), y1 = lds_u8(sym_saddr + s1), y2 = lds_u8(sym_saddr + s2), y3 = lds_u8(sym_saddr + s3); \
            const uint32_t n0 = e0 >> 12, n1 = e1 >> 12, n2 = e2 >> 12, n3 = e3 >> 12;                                          \
            const uint32_t n23 = n2 + n3, nbs = n0 + n1 + n23;                                                                  \
            const uint32_t incl = warp_incl_add_pred(nbs);                                                                      \
            uint32_t lo, hi;                                                                                                    \
            ring_lohi(cur - incl, lo, hi);              /* state 4l's bits are the uppermost of the lane's window */            \
            const uint32_t tot = __shfl_sync(FULL, incl, 31);                                                                   \
            if (tot > cur - floor_bits) { bad = 1; break; }                                                                     \
            const uint32_t w2 = __funnelshift_r(lo, hi, n3), w1 = __funnelshift_r(lo, hi, n23);                                 \
            const uint32_t w0 = __funnelshift_r(w1, hi >> n23, n1);                                                             \
            s0 = (e0 & 0xfffu) | (w0 & ~(0xffffffffu << n0));                                                                   \
            s1 = (e1 & 0xfffu) | (w1 & ~(0xffffffffu << n1));                                                                   \
            s2 = (e2 & 0xfffu) | (w2 & ~(0xffffffffu << n2));                                                                   \
            s3 = (e3 & 0xfffu) | (lo & ~(0xffffffffu << n3));                                                                   \
            const uint32_t sy = __byte_perm(__byte_perm(y0, y1, 0x0040), __byte_perm(y2, y3, 0x0040), 0x5410);                  \
            STORE;                                                                                                              \
            cur -= tot;                                                                                                         \
        }
        if (out_aligned) {
            uint32_t *const ow = reinterpret_cast<uint32_t *>(out);          // warp uniform; the index stays 32 bits wide
            for (; i0 + 128 <= body; i0 += 128) DEC128C_ROUND(ow[(i0 >> 2) + (uint32_t)lane] = sy)
        } else {
            for (; i0 + 128 <= body; i0 += 128)
                DEC128C_ROUND(out[i0 + 4 * lane] = (uint8_t)sy; out[i0 + 4 * lane + 1] = (uint8_t)(sy >> 8);
                              out[i0 + 4 * lane + 2] = (uint8_t)(sy >> 16); out[i0 + 4 * lane + 3] = (uint8_t)(sy >> 24))
        }
#undef DEC128C_ROUND
        if (!bad && i0 < body) {                            // last partial round
            stage();
            uint32_t ia = i0 + 4 * lane;
            uint32_t e0 = tab[s0], e1 = tab[s1], e2 = tab[s2], e3 = tab[s3];
            uint32_t n0 = ia < body ? (e0 >> 12) : 0u, n1 = ia + 1 < body ? (e1 >> 12) : 0u;
            uint32_t n2 = ia + 2 < body ? (e2 >> 12) : 0u, n3 = ia + 3 < body ? (e3 >> 12) : 0u;
            uint32_t n23 = n2 + n3, n123 = n1 + n23, nbs = n0 + n123;
            uint32_t incl = warp_incl_add_pred(nbs);
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) bad = 1;
            else {
                uint64_t w = ring_bits64(cur - incl);
                if (ia < body) { out[ia] = sym[s0]; s0 = (e0 & 0xfffu) | ((uint32_t)(w >> n123) & ~(0xffffffffu << n0)); }
                if (ia + 1 < body) { out[ia + 1] = sym[s1]; s1 = (e1 & 0xfffu) | ((uint32_t)(w >> n23) & ~(0xffffffffu << n1)); }
                if (ia + 2 < body) { out[ia + 2] = sym[s2]; s2 = (e2 & 0xfffu) | ((uint32_t)(w >> n3) & ~(0xffffffffu << n2)); }
                if (ia + 3 < body) { out[ia + 3] = sym[s3]; s3 = (e3 & 0xfffu) | ((uint32_t)w & ~(0xffffffffu << n3)); }
                cur -= tot;
            }
        }
        if (!bad) {                                         // Decoder::finish, fse.rs:383-385: i in [body, bn), state i % 128
            out[body + ((4 * lane - body) & 127)] = sym[s0];
            out[body + ((4 * lane + 1 - body) & 127)] = sym[s1];
            out[body + ((4 * lane + 2 - body) & 127)] = sym[s2];
            out[body + ((4 * lane + 3 - body) & 127)] = sym[s3];
        }
#if FSE_DEC_TMA
        if (pending) {                                      // never leave a copy in flight into memory the next block reuses
            if (!mbar_wait(bar, par)) tma_ok = 0;
            par ^= 1;
        }
        if (!tma_ok) bad = 1;
#endif
        cur -= floor_bits;
        if (bad || cur != 0) st = ST_LENGTH;
        if (lane == 0) a.status[b] = st;
    }
}

}  // namespace fsed
