// Image pyramid: chained bilinear down-scaling fused with the 7x7 Gaussian blur of every level.
//
// Replaces CpuImagePyramid::update (image_pyramid.cpp:68-86): level L = cv::resize(level L-1,
// INTER_LINEAR), blurred L = cv::GaussianBlur(level L, 7x7, sigma 2, BORDER_REFLECT_101).  Both
// OpenCV primitives are integer fixed-point on 8-bit data, so the kernel is bit-exact:
//   resize : 11-bit coefficient taps (host tables, tables.cpp), int32 row sums,
//            out = (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2
//   blur   : kernel {18,34,48,56,48,34,18}/256, horizontal 8.8 sums (u16), vertical 16.16 sums,
//            out = (v + 32768) >> 16, reflect-101 borders applied to the RESIZED pixels.
//
// One CTA produces a 64x32 tile of level L: the source tile of level L-1 is staged in shared
// memory, the resized tile plus a 3-px halo is produced into shared memory (and its interior
// written to the pyramid plane), halo pixels outside the image are mirrored in place, then the
// separable blur runs from shared memory and the blurred tile is written with 32-bit stores.
// Level L-1 is read once from HBM/L2 and both planes of level L are written once.
#include "ctx.h"
#include "tma.cuh"

namespace sg {

constexpr int TW = 64, TH = 32;          // output tile
constexpr int RW = TW + 8, RH = TH + 6;  // resized tile incl. halo; column 0 <-> x0-4 (word aligned)
constexpr int PYR_THREADS = 256;

struct PyrArgs {
    const uint8_t *src;   // level L-1 (or level L itself when !RESIZE)
    int sw, sh, spitch;
    unsigned long long sstride;
    uint8_t *dst;         // pyramid plane of level L (RESIZE only)
    uint8_t *blur;        // blurred plane of level L
    int w, h, pitch;
    unsigned long long dstride;
    const ResizeTap *xtab, *ytab;
    const uint4 *xtile, *ytile;  // fast kernel: per-tile tap tables (pyramid_tile_tables)
    int src_tile_w, src_tile_h;  // smem extent of the source tile (bytes per row multiple of 4)
    int area2x;
    int f0;               // first frame of the launch
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

template <bool RESIZE>
__global__ void __launch_bounds__(PYR_THREADS) pyr_level_kernel(const PyrArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *R = smem;                                        // [RH][RW]
    uint16_t *Hs = reinterpret_cast<uint16_t *>(smem + RH * RW);  // [RH][TW]
    uint8_t *S = smem + RH * RW + RH * TW * 2;                // [src_tile_h][src_tile_w]  (RESIZE)
    ResizeTap *xt = reinterpret_cast<ResizeTap *>(S + a.src_tile_h * a.src_tile_w);  // [TW+6]
    ResizeTap *yt = xt + (TW + 6);                                                  // [TH+6]

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, f = blockIdx.z + a.f0;
    const int tw = min(TW, a.w - x0), th = min(TH, a.h - y0);
    // in-image part of the halo window, [xlo, xhi) x [ylo, yhi)
    const int xlo = max(x0 - 3, 0), xhi = min(x0 + tw + 3, a.w);
    const int ylo = max(y0 - 3, 0), yhi = min(y0 + th + 3, a.h);
    const uint8_t *src = a.src + (size_t)f * a.sstride;

    if (RESIZE) {
        // ---- stage the source tile and the taps ------------------------------------------------
        const int sx_lo = a.xtab[xlo].s0 & ~3;
        const int sx_hi = a.area2x ? min(2 * (xhi - 1) + 1, a.sw - 1) : a.xtab[xhi - 1].s1;
        const int sy_lo = a.area2x ? 2 * ylo : a.ytab[ylo].s0;
        const int sy_hi = a.area2x ? min(2 * (yhi - 1) + 1, a.sh - 1) : a.ytab[yhi - 1].s1;
        const int nwords = ((sx_hi - sx_lo) >> 2) + 1, nrows = sy_hi - sy_lo + 1;
        const int sp = a.src_tile_w;
        for (int i = tid; i < nwords * nrows; i += PYR_THREADS) {
            const int r = i / nwords, wd = i - r * nwords;
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(
                src + (size_t)(sy_lo + r) * a.spitch + sx_lo + 4 * wd));
            *reinterpret_cast<uint32_t *>(S + r * sp + 4 * wd) = v;
        }
        for (int i = tid; i < (xhi - xlo); i += PYR_THREADS) {
            ResizeTap t = a.xtab[xlo + i];
            t.s0 -= sx_lo; t.s1 -= sx_lo;
            xt[i] = t;
        }
        for (int i = tid; i < (yhi - ylo); i += PYR_THREADS) {
            ResizeTap t = a.ytab[ylo + i];
            t.s0 -= sy_lo; t.s1 -= sy_lo;
            yt[i] = t;
        }
        __syncthreads();
        // ---- resized pixels of the in-image window -> R ----------------------------------------
        const int nx = xhi - xlo, ny = yhi - ylo;
        for (int i = tid; i < nx * ny; i += PYR_THREADS) {
            const int ry = i / nx, rx = i - ry * nx;
            int out;
            if (a.area2x) {
                const int sx = 2 * (xlo + rx) - sx_lo, sy = 2 * (ylo + ry) - sy_lo;
                out = (S[sy * sp + sx] + S[sy * sp + sx + 1] + S[(sy + 1) * sp + sx] + S[(sy + 1) * sp + sx + 1] + 2) >> 2;
            } else {
                const ResizeTap tx = xt[rx], ty = yt[ry];
                const uint8_t *r0 = S + ty.s0 * sp, *r1 = S + ty.s1 * sp;
                const int h0 = r0[tx.s0] * tx.a0 + r0[tx.s1] * tx.a1;
                const int h1 = r1[tx.s0] * tx.a0 + r1[tx.s1] * tx.a1;
                out = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
                out = min(max(out, 0), 255);
            }
            R[(ylo + ry - (y0 - 3)) * RW + (xlo + rx - (x0 - 4))] = (uint8_t)out;
        }
    } else {
        // ---- blur only: the window is read straight from the plane (aligned words) --------------
        const int ny = yhi - ylo;
        for (int i = tid; i < ny * (RW / 4); i += PYR_THREADS) {
            const int ry = i / (RW / 4), wd = i - ry * (RW / 4);
            const int x = x0 - 4 + 4 * wd;
            if (x >= 0 && x < a.spitch) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)(ylo + ry) * a.spitch + x));
                *reinterpret_cast<uint32_t *>(R + (ylo + ry - (y0 - 3)) * RW + 4 * wd) = v;
            }
        }
    }
    __syncthreads();

    // ---- pyramid plane: interior of R, 32-bit stores ----------------------------------------------
    if (RESIZE) {
        uint8_t *dst = a.dst + (size_t)f * a.dstride;
        for (int i = tid; i < th * (TW / 4); i += PYR_THREADS) {
            const int r = i / (TW / 4), wd = i - r * (TW / 4);
            if (4 * wd < tw) {
                const uint32_t v = *reinterpret_cast<const uint32_t *>(R + (r + 3) * RW + 4 + 4 * wd);
                *reinterpret_cast<uint32_t *>(dst + (size_t)(y0 + r) * a.pitch + x0 + 4 * wd) = v;
            }
        }
    }
    // ---- reflect-101: halo entries outside the image mirror resized pixels inside it ------------
    if (x0 == 0 || y0 == 0 || x0 + tw + 3 > a.w || y0 + th + 3 > a.h) {
        const int wx = tw + 6, wy = th + 6;
        for (int i = tid; i < wx * wy; i += PYR_THREADS) {
            const int ry = i / wx, rx = i - ry * wx;
            const int x = x0 - 3 + rx, y = y0 - 3 + ry;
            if (x < 0 || x >= a.w || y < 0 || y >= a.h) {
                const int mx = reflect101(x, a.w), my = reflect101(y, a.h);
                R[ry * RW + rx + 1] = R[(my - (y0 - 3)) * RW + (mx - (x0 - 4))];
            }
        }
        __syncthreads();
    }

    // ---- horizontal pass: 4 outputs per thread, two dp4a each ------------------------------------
    {
        const uint32_t k0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);   // taps 0..3
        const uint32_t k1 = 48u | (34u << 8) | (18u << 16);                 // taps 4..6
        for (int i = tid; i < (th + 6) * (TW / 4); i += PYR_THREADS) {
            const int ry = i / (TW / 4), g = i - ry * (TW / 4);
            const uint32_t *row = reinterpret_cast<const uint32_t *>(R + ry * RW) + g;
            const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];   // columns 4g .. 4g+11 of R
            // output c = 4g + j uses R columns c+1 .. c+7
            uint32_t o[4];
            o[0] = __dp4a(__byte_perm(w0, w1, 0x4321), k0, __dp4a(__byte_perm(w1, w2, 0x4321), k1, 0u));
            o[1] = __dp4a(__byte_perm(w0, w1, 0x5432), k0, __dp4a(__byte_perm(w1, w2, 0x5432), k1, 0u));
            o[2] = __dp4a(__byte_perm(w0, w1, 0x6543), k0, __dp4a(__byte_perm(w1, w2, 0x6543), k1, 0u));
            o[3] = __dp4a(w1, k0, __dp4a(w2, k1, 0u));
            uint2 st;
            st.x = o[0] | (o[1] << 16);
            st.y = o[2] | (o[3] << 16);
            *reinterpret_cast<uint2 *>(Hs + ry * TW + 4 * g) = st;
        }
    }
    __syncthreads();

    // ---- vertical pass: 4 columns x 2 rows per thread, 32-bit stores ------------------------------
    {
        uint8_t *bl = a.blur + (size_t)f * a.dstride;
        const int g = tid & 15, rp = tid >> 4;   // 16 column groups x 16 row pairs
        const int r0 = 2 * rp;
        if (r0 < th && 4 * g < tw) {
            uint32_t hv[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint2 v = *reinterpret_cast<const uint2 *>(Hs + (r0 + j) * TW + 4 * g);
                hv[j][0] = v.x & 0xffffu; hv[j][1] = v.x >> 16;
                hv[j][2] = v.y & 0xffffu; hv[j][3] = v.y >> 16;
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (r0 + rr < th) {
                    uint32_t packed = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t v = 18u * (hv[rr][c] + hv[rr + 6][c]) + 34u * (hv[rr + 1][c] + hv[rr + 5][c])
                                           + 48u * (hv[rr + 2][c] + hv[rr + 4][c]) + 56u * hv[rr + 3][c];
                        packed |= ((v + 32768u) >> 16) << (8 * c);
                    }
                    *reinterpret_cast<uint32_t *>(bl + (size_t)(y0 + r0 + rr) * a.pitch + x0 + 4 * g) = packed;
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Fast path (scale factor <= 2, no INTER_AREA switch): same arithmetic, restructured for the integer
// pipes, 64 x 64 output tiles (halo work 70/64 instead of 38/32, half the per-CTA fixed cost per pixel):
//   resize    : a thread owns one PAIR of window columns and walks down 9 (8) window rows of its warp, so the
//               column taps (16-bit coefficient pair, byte selectors into an aligned 8-byte source window) are
//               loop invariants and the 32 lanes of an instruction touch ONE window row (no bank conflicts); the
//               last 4 of the 36 column pairs run as an 8-row x 4-pair tail.  Per row: 4 LDS, 2 PRMT, 4 IDP.2A
//               (S0*a0 + S1*a1 of both columns and both source rows), then the two truncating vertical products.
//               Out-of-image window entries use clamped taps (overwritten by the mirror pass): no guards in the loop
//   blur H    : IDP.4A on byte windows (two per output), two rows per item, stored as vertical u16
//               pairs Hp[r/2][c] = H[r][c] | H[r+1][c] << 16
//   blur V    : IDP.2A on the vertical pairs (4 per output, rounding constant in the accumulator),
//               4 columns x 4 rows per thread from 5 LDS.128, 32-bit stores
// -------------------------------------------------------------------------------------------------
// R geometry of the fast kernel.  RESIZE: R is computed, pitch 80, column 4 <-> x0.  Blur only: R is the
// TMA box itself; TMA needs the innermost start coordinate to be a multiple of 16 BYTES (anything else
// raises "illegal instruction" on sm_100a -- measured), so the box starts at x0 - 16: pitch 96, column 16 <-> x0.
constexpr int FTH = 64, FRH = FTH + 6;       // output tile rows, window rows (incl. the 3-px blur halo)
constexpr int FR_BYTES = 6784;               // >= FRH * 96, multiple of 128
constexpr int COLP = 36;                     // window column pairs (72 columns: x0 - 4 .. x0 + 67)
constexpr int TAILP = COLP - 32;             // column pairs beyond a warp's 32 lanes
constexpr int MAIN_ROWS = 9;                 // warps 0..5: 9 window rows each, warps 6 and 7: 8 (54 + 16 = 70)
static_assert(TAILP == 4 && PYR_THREADS == 256 && 6 * MAIN_ROWS + 2 * (MAIN_ROWS - 1) == FRH, "resize work split");
template <bool RESIZE> struct RGeom { static constexpr int FW = RESIZE ? 80 : 96, X0 = RESIZE ? 4 : 16; };
struct XTap { uint32_t coef; int32_t s0; };          // a0 | a1 << 16, source column relative to the tile
struct YTap { uint32_t o0, o1, b0s, b1s; };          // byte offsets of the two source rows in S, coefficients << 12

template <bool RESIZE>
__global__ void __launch_bounds__(PYR_THREADS, 8)
pyr_fast_kernel(const PyrArgs a, const __grid_constant__ CUtensorMap tmap) {
    // [TMA destination: source tile S (RESIZE) or R itself] [R] [Hp] [taps of the 4 tail column pairs] [yt] [mbarrier]
    extern __shared__ __align__(128) uint8_t smem[];
    const int s_bytes = RESIZE ? (((a.src_tile_h + 1) * a.src_tile_w + 127) & ~127) : 0;
    uint8_t *S = smem;                                                   // [src_tile_h][src_tile_w] (RESIZE)
    uint8_t *R = smem + s_bytes;                                         // [FRH][FW]
    uint32_t *Hp = reinterpret_cast<uint32_t *>(R + FR_BYTES);           // [FRH/2][TW]
    uint4 *xtail = reinterpret_cast<uint4 *>(R + FR_BYTES + (FRH / 2) * TW * 4);   // [TAILP] (576 bytes reserved)
    YTap *yt = reinterpret_cast<YTap *>(xtail + 2 * COLP * sizeof(XTap) / sizeof(uint4));   // [FRH + 2]
    uint64_t *bar = reinterpret_cast<uint64_t *>(yt + FRH + 2);

    constexpr int FW = RGeom<RESIZE>::FW, X0 = RGeom<RESIZE>::X0;   // R pitch; R column of x0
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * FTH, f = blockIdx.z + a.f0;
    const int tw = min(TW, a.w - x0), th = min(FTH, a.h - y0);
    const int xlo = max(x0 - 3, 0), xhi = min(x0 + tw + 3, a.w);
    const int ylo = max(y0 - 3, 0), yhi = min(y0 + th + 3, a.h);

    if (RESIZE) {
        // one thread pulls the source tile with TMA while the others stage the taps
        if (tid == 0) {
            const int sx_lo = a.xtab[xlo].s0 & ~15, sy_lo = a.ytab[ylo].s0;   // 16-byte aligned box origin
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (uint32_t)(a.src_tile_w * a.src_tile_h));
            tma_load_3d(S, &tmap, sx_lo, sy_lo, f, bar);
        }
        // taps: the window rows' entries go to shared memory (read once per row by every thread), the thread's own column
        // pair comes straight from the per-tile table (pyramid_tile_tables: nothing here depends on the frame)
        if (tid >= 128 && tid < 128 + FRH) reinterpret_cast<uint4 *>(yt)[tid - 128] = __ldg(a.ytile + blockIdx.y * FRH + (tid - 128));
        if (tid >= 128 + FRH && tid < 128 + FRH + TAILP)
            xtail[tid - 128 - FRH] = __ldg(a.xtile + blockIdx.x * COLP + 32 + (tid - 128 - FRH));
        // Work split: a warp never spans two window rows in one instruction of the main pass (lane = column pair 0..31 of ONE
        // row, so the gathered source words and the 16-bit stores of a warp fall into distinct banks); the remaining
        // TAILP column pairs are done 8 rows x TAILP pairs per warp instruction.  Per warp: 9 (8) main rows + 1 (2) tail passes.
        const int lane = tid & 31, wp = tid >> 5;
        // {coef a, coef b, byte offset of the aligned 8-byte source window, PRMT selector gathering {S[a], S[a+1], S[b], S[b+1]}}
        const uint4 tx = __ldg(a.xtile + blockIdx.x * COLP + lane);
        __syncthreads();          // taps staged, barrier initialised
        const uint32_t s_base = smem_u32(S), r_base = smem_u32(R), yt_base = smem_u32(yt);
        auto resize_px = [&](const uint4 &t, int row) {     // window row `row`, the column pair t describes -> two R bytes
            const uint4 ty = lds128(yt_base + (unsigned)row * (unsigned)sizeof(YTap));
            const uint32_t a0 = s_base + t.z + ty.x, a1 = s_base + t.z + ty.y;
            const uint32_t g0 = __byte_perm(lds32(a0), lds32(a0 + 4), t.w), g1 = __byte_perm(lds32(a1), lds32(a1 + 4), t.w);
            const unsigned ha0 = __dp2a_lo(t.x, g0, 0u), hb0 = __dp2a_hi(t.y, g0, 0u);
            const unsigned ha1 = __dp2a_lo(t.x, g1, 0u), hb1 = __dp2a_hi(t.y, g1, 0u);
            // ((b * (h >> 4)) >> 16) == umulhi(b << 12, h & ~15): one LOP3 + one IMAD.HI per product
            const unsigned va = (__umulhi(ty.w, ha1 & ~15u) + __umulhi(ty.z, ha0 & ~15u) + 2u) >> 2;
            const unsigned vb = (__umulhi(ty.w, hb1 & ~15u) + __umulhi(ty.z, hb0 & ~15u) + 2u) >> 2;
            return va | (vb << 8);
        };
        mbar_wait(bar, 0);        // source tile landed
        // partial tiles (right / bottom image edge): only the window columns / rows the blur of the tile reads
        const int rows_needed = min(FRH, th + 6);
        {
            const int row0 = wp < 6 ? MAIN_ROWS * wp : 6 * MAIN_ROWS + (MAIN_ROWS - 1) * (wp - 6);
            const int n = min(wp < 6 ? MAIN_ROWS : MAIN_ROWS - 1, rows_needed - row0);
            if (2 * lane < tw + 8) {
                const uint32_t r_addr = r_base + 2u * lane + (unsigned)row0 * FW;
                if (n >= MAIN_ROWS - 1) {
#pragma unroll
                    for (int k = 0; k < MAIN_ROWS - 1; ++k) sts16(r_addr + k * FW, resize_px(tx, row0 + k));
                    if (n == MAIN_ROWS) sts16(r_addr + (MAIN_ROWS - 1) * FW, resize_px(tx, row0 + MAIN_ROWS - 1));
                } else {
                    for (int k = 0; k < n; ++k) sts16(r_addr + k * FW, resize_px(tx, row0 + k));
                }
            }
        }
        {
            const int cq = lane & (TAILP - 1), rq = lane >> 2;
            if (2 * (32 + cq) < tw + 8) {
                const uint4 tt = lds128(smem_u32(xtail) + (unsigned)cq * 16u);
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {                          // warp 6 also takes window rows 64 .. 69
                    if (pass == 1 && wp != 6) break;
                    const int row = (pass ? 64 : 8 * wp) + rq;
                    if (row < rows_needed) sts16(r_base + 2u * (32 + cq) + (unsigned)row * FW, resize_px(tt, row));
                }
            }
        }
    } else {
        // blur only: the 96 x 70 window of the plane goes straight into R (zero outside the image)
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (uint32_t)(FRH * FW));
            tma_load_3d(R, &tmap, x0 - X0, y0 - 3, f, bar);
        }
        __syncthreads();
        mbar_wait(bar, 0);
    }
    __syncthreads();

    // (the pyramid plane itself -- the interior of R -- is stored by the vertical blur pass below: same rows, same words,
    //  same address arithmetic as the blurred plane)
    // ---- reflect-101: halo entries outside the image mirror resized pixels inside it ------------
    // Rows first (whole words of the up to three rows above / below the image, copied from their mirror rows), then,
    // behind a barrier, the up to three columns left / right of the image for EVERY window row (the halo rows now hold
    // their mirrors, so a corner gets the pixel mirrored in both directions).  One thread per word / per row and side.
    {
        const bool left = x0 == 0, right = x0 + tw + 3 > a.w, top = y0 == 0, bottom = y0 + th + 3 > a.h;
        if (top || bottom) {
            constexpr int WORDS = 18;                               // window columns x0 - 4 .. x0 + 67
            if (tid < 6 * WORDS) {
                const int k = tid / WORDS, wd = tid - k * WORDS;    // k < 3: row -1 - k ; else row a.h + (k - 3)
                const int y = k < 3 ? -1 - k : a.h + k - 3;
                const int ry = y - (y0 - 3);
                if ((k < 3 ? top : bottom) && ry < FRH) {
                    const int my = k < 3 ? -y : 2 * (a.h - 1) - y;   // reflect-101; inside the window: |y - my| <= 6
                    uint32_t *row = reinterpret_cast<uint32_t *>(R + ry * FW + X0 - 4) + wd;
                    *row = *(reinterpret_cast<const uint32_t *>(R + (my - (y0 - 3)) * FW + X0 - 4) + wd);
                }
            }
            if (left || right) __syncthreads();
        }
        if (left || right) {
            const int ry = tid & 127, side = tid >> 7;              // threads 0..69: left strip, 128..197: right strip
            if (ry < FRH && (side == 0 ? left : right)) {
                uint8_t *row = R + ry * FW + X0 - x0;               // row[x] = window pixel x
                if (side == 0) {
                    row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3];
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int x = a.w + k;
                        if (x - x0 < TW + 4) row[x] = row[2 * (a.w - 1) - x];
                    }
                }
            }
        }
        if (left || right || top || bottom) __syncthreads();
    }

    // ---- horizontal pass: 4 columns x 2 rows per item, stored as vertical u16 pairs ---------------
    // Output c = 4g + j is sum_i k[i] * R[c + 1 + i] over the 12 window bytes w0 w1 w2 = R[4g .. 4g + 11]:
    // the kernel is shifted inside the coefficient words instead of shifting the data, so the pass is
    // IDP.4A only (10 per 4 outputs, FMA pipe) with no PRMT on the ALU pipe.
    {
        constexpr uint32_t K0 = 18, K1 = 34, K2 = 48, K3 = 56;                      // k[0..3]; k[4..6] = k[2..0]
        constexpr uint32_t A0 = (K0 << 8) | (K1 << 16) | (K2 << 24);                // j = 0: w0 bytes 1..3
        constexpr uint32_t A1 = K3 | (K2 << 8) | (K1 << 16) | (K0 << 24);           //        w1 bytes 0..3
        constexpr uint32_t B0 = (K0 << 16) | (K1 << 24);                            // j = 1: w0 bytes 2..3
        constexpr uint32_t B1 = K2 | (K3 << 8) | (K2 << 16) | (K1 << 24);           //        w1
        constexpr uint32_t B2 = K0;                                                 //        w2 byte 0
        constexpr uint32_t C0 = K0 << 24;                                           // j = 2: w0 byte 3
        constexpr uint32_t C1 = K1 | (K2 << 8) | (K3 << 16) | (K2 << 24);           //        w1
        constexpr uint32_t C2 = K1 | (K0 << 8);                                     //        w2 bytes 0..1
        constexpr uint32_t D1 = K0 | (K1 << 8) | (K2 << 16) | (K3 << 24);           // j = 3: w1
        constexpr uint32_t D2 = K2 | (K1 << 8) | (K0 << 16);                        //        w2 bytes 0..2
        const int g = tid & 15;
        const int rp_end = min(FRH / 2, (th + 7) >> 1);     // partial tiles: rows / columns of the tile only
        // the two half warps take row pairs rp and rp + 2 (bits 0 and 1 of tid >> 4 swapped): their R rows are 4 apart,
        // 4 * FW / 4 = 80 or 96 words = 16 banks mod 32 for FW = 80 (0 for 96, where the 24-word rows already interleave)
        const int q = tid >> 4;
        const int rp_first = RESIZE ? ((q & ~3) | ((q & 1) << 1) | ((q >> 1) & 1)) : q;
        for (int rp = rp_first; rp < rp_end && 4 * g < tw; rp += PYR_THREADS / 16) {
            uint32_t o[2][4];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint32_t *row = reinterpret_cast<const uint32_t *>(R + (2 * rp + rr) * FW + X0 - 4) + g;
                const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];   // window columns 4g .. 4g+11 (x0-4+4g ..)
                o[rr][0] = __dp4a(w1, A1, __dp4a(w0, A0, 0u));
                o[rr][1] = __dp4a(w2, B2, __dp4a(w1, B1, __dp4a(w0, B0, 0u)));
                o[rr][2] = __dp4a(w2, C2, __dp4a(w1, C1, __dp4a(w0, C0, 0u)));
                o[rr][3] = __dp4a(w2, D2, __dp4a(w1, D1, 0u));
            }
            // H[r] | H[r+1] << 16 as a multiply-add (FMA pipe); both halves are below 65536
            *reinterpret_cast<uint4 *>(Hp + rp * TW + 4 * g) =
                make_uint4(o[1][0] * 65536u + o[0][0], o[1][1] * 65536u + o[0][1],
                           o[1][2] * 65536u + o[0][2], o[1][3] * 65536u + o[0][3]);
        }
    }
    __syncthreads();

    // ---- vertical pass: 4 columns x 4 rows per thread; out row r uses H rows r .. r+6 --------------
    {
        const int g = tid & 15, strip = tid >> 4;   // 16 column groups x 16 strips of 2 row pairs
        const int r0 = 4 * strip;
        if (r0 < th && 4 * g < tw) {
            uint8_t *bl = a.blur + (size_t)f * a.dstride + (size_t)(y0 + r0) * a.pitch + x0 + 4 * g;
            const ptrdiff_t plane_delta = RESIZE ? a.dst - a.blur : 0;       // same layout, other plane
            const uint32_t r_word = smem_u32(R) + (unsigned)((r0 + 3) * FW + X0 + 4 * g);
            uint4 P[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) P[j] = *reinterpret_cast<const uint4 *>(Hp + (2 * strip + j) * TW + 4 * g);
            const uint32_t kA = 18u | (34u << 8) | (48u << 16) | (56u << 24);   // k0 k1 | k2 k3
            const uint32_t kB = 48u | (34u << 8);                               // k4 k5
            const uint32_t kC = 34u | (48u << 8) | (56u << 16) | (48u << 24);   // k1 k2 | k3 k4
            const uint32_t kD = 34u | (18u << 8);                               // k5 k6
            const uint32_t kE = 18u;                                            // k6 on the low half
            const uint32_t kF = 18u << 8;                                       // k0 on the high half
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t p0[4] = {P[h].x, P[h].y, P[h].z, P[h].w}, p1[4] = {P[h + 1].x, P[h + 1].y, P[h + 1].z, P[h + 1].w};
                const uint32_t p2[4] = {P[h + 2].x, P[h + 2].y, P[h + 2].z, P[h + 2].w}, p3[4] = {P[h + 3].x, P[h + 3].y, P[h + 3].z, P[h + 3].w};
                uint32_t ve[4], vo[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // even row: rows r..r+5 are the pairs p0 p1 p2, row r+6 is the low half of p3
                    uint32_t e = __dp2a_lo(p0[c], kA, 32768u);
                    e = __dp2a_hi(p1[c], kA, e);
                    e = __dp2a_lo(p2[c], kB, e);
                    ve[c] = __dp2a_lo(p3[c], kE, e);
                    // odd row: row r+1 is the high half of p0, rows r+2..r+7 are p1 p2 p3
                    uint32_t o = __dp2a_lo(p0[c], kF, 32768u);
                    o = __dp2a_lo(p1[c], kC, o);
                    o = __dp2a_hi(p2[c], kC, o);
                    vo[c] = __dp2a_lo(p3[c], kD, o);
                }
                // byte 2 of every sum (v >> 16 fits 8 bits): two PRMTs gather four of them
                const uint32_t even = __byte_perm(__byte_perm(ve[0], ve[1], 0x0062), __byte_perm(ve[2], ve[3], 0x0062), 0x5410);
                const uint32_t odd = __byte_perm(__byte_perm(vo[0], vo[1], 0x0062), __byte_perm(vo[2], vo[3], 0x0062), 0x5410);
                const int r = r0 + 2 * h;
                if (r < th) {
                    uint8_t *o = bl + (size_t)(2 * h) * a.pitch;
                    *reinterpret_cast<uint32_t *>(o) = even;
                    if (RESIZE) *reinterpret_cast<uint32_t *>(o + plane_delta) = lds32(r_word + (unsigned)(2 * h) * FW);
                }
                if (r + 1 < th) {
                    uint8_t *o = bl + (size_t)(2 * h + 1) * a.pitch;
                    *reinterpret_cast<uint32_t *>(o) = odd;
                    if (RESIZE) *reinterpret_cast<uint32_t *>(o + plane_delta) = lds32(r_word + (unsigned)(2 * h + 1) * FW);
                }
            }
        }
    }
}

static size_t pyr_fast_smem_bytes(const PyrArgs &a, bool resize) {
    size_t b = FR_BYTES + (FRH / 2) * TW * 4 + 2 * COLP * sizeof(XTap) + (FRH + 2) * sizeof(YTap) + 16;
    if (resize) b += (((size_t)(a.src_tile_h + 1) * a.src_tile_w + 127) & ~(size_t)127);
    return b;
}

// Source rows one 64-row tile of the fast kernel needs (TMA box height), host, at sg_create.
int pyramid_fast_source_rows(const std::vector<ResizeTap> &yt, int h) {
    int mh = 0;
    for (int y0 = 0; y0 < h; y0 += FTH) {
        const int th = std::min(FTH, h - y0), ylo = std::max(y0 - 3, 0), yhi = std::min(y0 + th + 3, h);
        mh = std::max(mh, yt[yhi - 1].s1 - yt[ylo].s0 + 1);
    }
    return mh;
}
int pyramid_fast_tile_rows() { return FTH; }

// Per-tile tap tables of the fast resize kernel (host, at sg_create): what the kernel used to derive per CTA from xtab /
// ytab.  xtile[tile column][COLP]: for the window column pair (x0 - 4 + 2cp, + 1), clamped to the in-image window like the
// kernel's halo handling: {a0 | a1 << 16 of column a, of column b, byte offset of the aligned 8-byte source window inside
// the TMA tile, PRMT selector gathering {S[a], S[a+1], S[b], S[b+1]} from it}.  ytile[tile row][FRH]: {byte offset of the
// two source rows in the tile, b0 << 12, b1 << 12}.  The TMA box origin (xtab[xlo].s0 & ~15, ytab[ylo].s0) stays in the kernel.
void pyramid_tile_tables(const std::vector<ResizeTap> &xt, const std::vector<ResizeTap> &yt, int w, int h, int src_pitch,
                         std::vector<uint4> &xtile, std::vector<uint4> &ytile) {
    xtile.clear(); ytile.clear();
    for (int x0 = 0; x0 < w; x0 += TW) {
        const int tw = std::min(TW, w - x0), xlo = std::max(x0 - 3, 0), xhi = std::min(x0 + tw + 3, w);
        const int sx_lo = xt[xlo].s0 & ~15;
        for (int cp = 0; cp < COLP; ++cp) {
            const ResizeTap ta = xt[std::min(std::max(x0 - 4 + 2 * cp, xlo), xhi - 1)];
            const ResizeTap tb = xt[std::min(std::max(x0 - 4 + 2 * cp + 1, xlo), xhi - 1)];
            const int sa = ta.s0 - sx_lo, sb = tb.s0 - sx_lo, base = sa >> 2;
            const unsigned oa = (unsigned)(sa & 3), ob = (unsigned)(sb - 4 * base);
            uint4 e;
            e.x = (uint32_t)(uint16_t)ta.a0 | ((uint32_t)(uint16_t)ta.a1 << 16);
            e.y = (uint32_t)(uint16_t)tb.a0 | ((uint32_t)(uint16_t)tb.a1 << 16);
            e.z = 4u * (unsigned)base;
            e.w = (oa * 0x11u + 0x10u) | ((ob * 0x11u + 0x10u) << 8);
            xtile.push_back(e);
        }
    }
    for (int y0 = 0; y0 < h; y0 += FTH) {
        const int th = std::min(FTH, h - y0), ylo = std::max(y0 - 3, 0), yhi = std::min(y0 + th + 3, h);
        const int sy_lo = yt[ylo].s0;
        for (int r = 0; r < FRH; ++r) {
            const ResizeTap t = yt[std::min(std::max(y0 - 3 + r, ylo), yhi - 1)];
            uint4 e;
            e.x = (uint32_t)((t.s0 - sy_lo) * src_pitch); e.y = (uint32_t)((t.s1 - sy_lo) * src_pitch);
            e.z = (uint32_t)(uint16_t)t.a0 << 12; e.w = (uint32_t)(uint16_t)t.a1 << 12;
            ytile.push_back(e);
        }
    }
}

static size_t pyr_smem_bytes(const PyrArgs &a, bool resize) {
    size_t b = RH * RW + RH * TW * 2;
    if (resize) b += (size_t)a.src_tile_h * a.src_tile_w + sizeof(ResizeTap) * (TW + 6 + TH + 6);
    return b;
}

// Largest source-tile extent over all tiles of a level (host, at sg_create).
void pyramid_source_extent(const std::vector<ResizeTap> &xt, const std::vector<ResizeTap> &yt, int w, int h,
                           bool area2x, int sw, int sh, int *tile_w, int *tile_h) {
    int mw = 0, mh = 0;
    for (int x0 = 0; x0 < w; x0 += TW) {
        const int tw = std::min(TW, w - x0), xlo = std::max(x0 - 3, 0), xhi = std::min(x0 + tw + 3, w);
        const int lo = (area2x ? 2 * xlo : xt[xlo].s0) & ~3;
        const int hi = area2x ? std::min(2 * (xhi - 1) + 1, sw - 1) : xt[xhi - 1].s1;
        mw = std::max(mw, (((hi - lo) >> 2) + 1) * 4);
    }
    for (int y0 = 0; y0 < h; y0 += TH) {
        const int th = std::min(TH, h - y0), ylo = std::max(y0 - 3, 0), yhi = std::min(y0 + th + 3, h);
        const int lo = area2x ? 2 * ylo : yt[ylo].s0;
        const int hi = area2x ? std::min(2 * (yhi - 1) + 1, sh - 1) : yt[yhi - 1].s1;
        mh = std::max(mh, hi - lo + 1);
    }
    *tile_w = mw;
    *tile_h = mh;
}

int launch_pyramid(sg_ctx *ctx, int n_frames) {
    const int levels = ctx->p.levels;
    mark(ctx, EV_PYR0, true);
    for (int l = 0; l < levels; ++l) {
        const Level &L = ctx->lv[l];
        PyrArgs a{};
        a.f0 = ctx->frame0;
        a.w = L.w; a.h = L.h; a.pitch = L.pitch; a.dstride = L.frame_stride;
        a.blur = L.blur;
        dim3 grid((L.w + TW - 1) / TW, (L.h + TH - 1) / TH, n_frames);
        const dim3 fgrid((L.w + TW - 1) / TW, (L.h + FTH - 1) / FTH, n_frames);
        if (l == 0) {
            a.src = ctx->level0; a.sw = L.w; a.sh = L.h; a.spitch = ctx->level0_pitch; a.sstride = ctx->level0_stride;
            pyr_fast_kernel<false><<<fgrid, PYR_THREADS, pyr_fast_smem_bytes(a, false), ctx->stream>>>(a, L.map_src);
        } else {
            const Level &P = ctx->lv[l - 1];
            a.src = l == 1 ? ctx->level0 : P.pyr;
            a.sw = P.w; a.sh = P.h;
            a.spitch = l == 1 ? ctx->level0_pitch : P.pitch;
            a.sstride = l == 1 ? ctx->level0_stride : P.frame_stride;
            a.dst = L.pyr;
            a.xtab = L.xtab; a.ytab = L.ytab;
            a.xtile = L.xtile; a.ytile = L.ytile;
            a.src_tile_w = L.src_tile_w; a.src_tile_h = L.src_tile_h;
            a.area2x = L.area2x ? 1 : 0;
            PyrArgs fa = a;
            fa.src_tile_w = L.tma_src_w; fa.src_tile_h = L.tma_src_h;
            const size_t fsmem = pyr_fast_smem_bytes(fa, true);
            if (L.fast_resize && fsmem <= 48 * 1024) {
                pyr_fast_kernel<true><<<fgrid, PYR_THREADS, fsmem, ctx->stream>>>(fa, L.map_src);
            } else {   // INTER_AREA switch (exact 2x) or a scale factor above 2: generic kernel
                const size_t smem = pyr_smem_bytes(a, true);
                if (smem > 48 * 1024)
                    SG_CUDA(ctx, cudaFuncSetAttribute(pyr_level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                pyr_level_kernel<true><<<grid, PYR_THREADS, smem, ctx->stream>>>(a);
            }
        }
        SG_LAUNCH_CHECK(ctx);
    }
    mark(ctx, EV_PYR1);
    return SG_OK;
}

}  // namespace sg
