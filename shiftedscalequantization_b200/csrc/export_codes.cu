// True-integer export of calibrated weights. The reference only ever holds dequantised fp32
// weights: the integer code is the intermediate `x_quant` of UniformAffineQuantizer.forward (quant/quant_layer.py:92-96),
// AdaRoundQuantizer.forward hard mode (quant/adaptive_rounding.py:50-58) and ChannelQuantMSE.forward
// (quant/channelQuantMSE.py:134-141). These kernels evaluate exactly that intermediate (same IEEE division, rint / floor,
// clamp) and store it bit-packed, and read it back into the dequantised weight `(q - zp) * delta [* in_scale]` the
// reference's forward returns — import(export(w)) is bit-identical to the hard forward.
//
// Packed layout: the tensor is [rows, k] (rows = output channels); every row starts on a byte boundary and takes
// ssq_packed_row_bytes(k, n_bits) bytes; codes are stored as u = q - qmin in `sbits` = smallest of {1,2,4,8} >= n_bits
// bits, element j of a row at bit (j % (8/sbits)) * sbits of byte j / (8/sbits) (little-endian within the byte).
//
// HBM-bound: (4 [+4 with alpha] + sbits/8) B/elem out, the reverse in. Vector path (rows whose length is a multiple of
// 32/sbits, channels a multiple of 4 long, 16-byte aligned): one thread = one float4, lanes cooperate on the code words.
#include "ssq_common.cuh"

namespace ssq {

__host__ __device__ __forceinline__ int storage_bits(int n_bits) { return n_bits <= 1 ? 1 : (n_bits <= 2 ? 2 : (n_bits <= 4 ? 4 : 8)); }

struct ExportArgs {
    const float* w; const float* alpha; const float* in_scale; const float* delta; const float* zp;
    uint8_t* packed;
    int64_t rows, k, inner, nchan, row_bytes;
    float qmin, qmax;
};

// q - qmin for one element (integer-valued float in [0, qmax-qmin])
__device__ __forceinline__ uint32_t code_of(float x, float a, bool has_alpha, const Recip& R, float zp, float qmin, float qmax) {
    const float t = div_exact(x, R);
    const float r = has_alpha ? (floorf(t) + (a >= 0.0f ? 1.0f : 0.0f)) : rintf(t);
    const float q = fminf(fmaxf(r + zp, qmin), qmax);
    return (uint32_t)(int)rintf(q - qmin);
}

// channel of flat element i: c = (i / inner) % nchan
__device__ __forceinline__ int64_t chan_of(int64_t i, int64_t inner, int64_t nchan) {
    if (((uint64_t)(i | inner | nchan) >> 32) == 0) return (int64_t)(((uint32_t)i / (uint32_t)inner) % (uint32_t)nchan);
    return (i / inner) % nchan;
}

// ---- vector path ---------------------------------------------------------------------------------------------
// One thread = one float4 of weights (4 codes = 4*SBITS bits), consecutive lanes = consecutive float4s, so every load
// instruction of a warp is one contiguous 512-byte run. G = 8/SBITS neighbouring lanes hold the pieces of one 32-bit
// word of codes: they are OR-ed together with a shuffle butterfly and lane 0 of the group stores the word.
// Needs k % (32/SBITS) == 0 (rows are whole words) and inner % 4 == 0 (a float4 never straddles two channels);
// tiles are address-ordered (ssq_common.cuh), the channel of a vector comes from TileWalk.
template <int SBITS, bool HAS_ALPHA>
__device__ __forceinline__ uint32_t pack4(const float4& x, const float4& al, float d, float zp, float qmin, float qmax) {
    const Recip R = make_recip(d);
    const float4 t = div4_exact(x, R);
    float4 r;
    if (HAS_ALPHA) {
        r.x = floorf(t.x) + (al.x >= 0.f ? 1.f : 0.f); r.y = floorf(t.y) + (al.y >= 0.f ? 1.f : 0.f);
        r.z = floorf(t.z) + (al.z >= 0.f ? 1.f : 0.f); r.w = floorf(t.w) + (al.w >= 0.f ? 1.f : 0.f);
    } else {
        r.x = rintf(t.x); r.y = rintf(t.y); r.z = rintf(t.z); r.w = rintf(t.w);
    }
    // q - qmin is a small non-negative integer held exactly in fp32
    const uint32_t u0 = __float2uint_rn(fminf(fmaxf(r.x + zp, qmin), qmax) - qmin);
    const uint32_t u1 = __float2uint_rn(fminf(fmaxf(r.y + zp, qmin), qmax) - qmin);
    const uint32_t u2 = __float2uint_rn(fminf(fmaxf(r.z + zp, qmin), qmax) - qmin);
    const uint32_t u3 = __float2uint_rn(fminf(fmaxf(r.w + zp, qmin), qmax) - qmin);
    return u0 | (u1 << SBITS) | (u2 << (2 * SBITS)) | (u3 << (3 * SBITS));
}

template <int SBITS, bool HAS_ALPHA>
__global__ void __launch_bounds__(SSQ_THREADS)
export_vec_kernel(ExportArgs a) {
    constexpr int G = 8 / SBITS;                 // lanes per output word: 8, 4, 2, 1
    constexpr int U = 4;                         // float4s per thread; one address-ordered tile of 256*U per CTA
    const int64_t total4 = (a.rows * a.k) >> 2;  // a multiple of G (rows are whole words), so a lane group is all-in or all-out
    const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int sh = (lane & (G - 1)) * 4 * SBITS;
    uint32_t* __restrict__ out = reinterpret_cast<uint32_t*>(a.packed);
    float4 x[U], al[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        x[u] = al[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total4) { x[u] = ld_stream4(a.w + i * 4); if (HAS_ALPHA) al[u] = ld_stream4(a.alpha + i * 4); }
    }
    TileWalk tw;
    tw.init((uint64_t)i0, (uint64_t)(a.inner >> 2), (uint64_t)a.nchan);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        uint32_t piece = 0;
        if (i < total4)
            piece = pack4<SBITS, HAS_ALPHA>(x[u], al[u], __ldg(a.delta + tw.c), __ldg(a.zp + tw.c), a.qmin, a.qmax) << sh;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) piece |= __shfl_xor_sync(0xffffffffu, piece, o);
        if (i < total4 && (lane & (G - 1)) == 0) out[i / G] = piece;
        tw.step(SSQ_THREADS);
    }
}

// ---- general path: one thread = one output byte (ragged rows: 9, 27, 147 ...; unaligned pointers) -----------------
template <bool HAS_ALPHA>
__global__ void __launch_bounds__(SSQ_THREADS)
export_byte_kernel(ExportArgs a, int sbits) {
    const int per = 8 / sbits;
    const int64_t total = a.rows * a.row_bytes;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t row = t / a.row_bytes, j0 = (t - row * a.row_bytes) * per;
        uint32_t byte = 0;
        for (int e = 0; e < per; ++e) {
            const int64_t j = j0 + e;
            if (j >= a.k) break;                 // tail of the row: padding bits stay zero
            const int64_t i = row * a.k + j;
            const int64_t c = chan_of(i, a.inner, a.nchan);
            float x = a.w[i];
            if (a.in_scale) x = __fdiv_rn(x, __ldg(a.in_scale + j));
            const Recip R = make_recip(__ldg(a.delta + c));
            byte |= code_of(x, HAS_ALPHA ? a.alpha[i] : 0.f, HAS_ALPHA, R, __ldg(a.zp + c), a.qmin, a.qmax) << (e * sbits);
        }
        a.packed[t] = (uint8_t)byte;
    }
}

struct ImportArgs {
    const uint8_t* packed; const float* in_scale; const float* delta; const float* zp; float* wq;
    int64_t rows, k, inner, nchan, row_bytes;
    float qmin;
};

__device__ __forceinline__ float dequant_of(uint32_t u, float qmin, float zp, float d) {
    return __fmul_rn(__fsub_rn((float)u + qmin, zp), d);        // (x_quant - zero_point) * delta
}

// one thread = one float4 of dequantised weights; the G lanes of a word all load it (one broadcast sector)
template <int SBITS>
__global__ void __launch_bounds__(SSQ_THREADS)
import_vec_kernel(ImportArgs a) {
    constexpr int G = 8 / SBITS;
    constexpr int U = 4;
    constexpr uint32_t MASK = (1u << SBITS) - 1u;
    const int64_t total4 = (a.rows * a.k) >> 2;
    const int64_t i0 = (int64_t)blockIdx.x * (SSQ_THREADS * U) + threadIdx.x;
    const uint32_t* __restrict__ in = reinterpret_cast<const uint32_t*>(a.packed);
    uint32_t bits[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        bits[u] = (i < total4) ? __ldg(in + i / G) >> ((uint32_t)(i & (G - 1)) * 4 * SBITS) : 0u;
    }
    TileWalk tw;
    tw.init((uint64_t)i0, (uint64_t)(a.inner >> 2), (uint64_t)a.nchan);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * SSQ_THREADS;
        if (i < total4) {
            const float d = __ldg(a.delta + tw.c), zp = __ldg(a.zp + tw.c);
            float4 y;
            y.x = dequant_of(bits[u] & MASK, a.qmin, zp, d);
            y.y = dequant_of((bits[u] >> SBITS) & MASK, a.qmin, zp, d);
            y.z = dequant_of((bits[u] >> (2 * SBITS)) & MASK, a.qmin, zp, d);
            y.w = dequant_of((bits[u] >> (3 * SBITS)) & MASK, a.qmin, zp, d);
            st_stream4(a.wq + i * 4, y);
        }
        tw.step(SSQ_THREADS);
    }
}

__global__ void __launch_bounds__(SSQ_THREADS)
import_elem_kernel(ImportArgs a, int sbits) {
    const int per = 8 / sbits;
    const uint32_t mask = (1u << sbits) - 1u;
    const int64_t total = a.rows * a.k;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t row = i / a.k, j = i - row * a.k;
        const uint32_t byte = a.packed[row * a.row_bytes + j / per];
        const int64_t c = chan_of(i, a.inner, a.nchan);
        float y = dequant_of((byte >> ((j % per) * sbits)) & mask, a.qmin, __ldg(a.zp + c), __ldg(a.delta + c));
        if (a.in_scale) y = __fmul_rn(y, __ldg(a.in_scale + j));
        a.wq[i] = y;
    }
}

static bool layout_ok(int64_t rows, int64_t k, int64_t inner, int64_t nchan) {
    if (rows < 0 || k < 0 || inner <= 0 || nchan <= 0) return false;
    const int64_t n = rows * k;
    return n % inner == 0 && (n / inner) % nchan == 0;
}

}  // namespace ssq

using namespace ssq;

extern "C" int64_t ssq_packed_row_bytes(int64_t k, int n_bits) {
    if (k < 0 || n_bits < 1 || n_bits > 8) return -1;
    return (k * storage_bits(n_bits) + 7) / 8;
}

extern "C" int ssq_export_codes(const float* w, const float* alpha, const float* in_scale, const float* delta,
                                const float* zero_point, uint8_t* packed, int64_t rows, int64_t k, int64_t inner,
                                int64_t nchan, float qmin, float qmax, int n_bits, void* stream) {
    if (rows == 0 || k == 0) return SSQ_OK;
    if (!w || !delta || !zero_point || !packed) return SSQ_ERR_NULL;
    if (n_bits < 1 || n_bits > 8 || !layout_ok(rows, k, inner, nchan)) return SSQ_ERR_SIZE;
    if (!(qmax >= qmin) || (qmax - qmin) > (float)((1 << n_bits) - 1)) return SSQ_ERR_SIZE;
    const int sbits = storage_bits(n_bits);
    ExportArgs a{w, alpha, in_scale, delta, zero_point, packed, rows, k, inner, nchan, ssq_packed_row_bytes(k, n_bits), qmin, qmax};
    cudaStream_t st = (cudaStream_t)stream;
    const int E = 32 / sbits;
    const bool vec = (k % E == 0) && (inner % 4 == 0) && !in_scale && aligned16(w) && (!alpha || aligned16(alpha)) &&
                     ((reinterpret_cast<uintptr_t>(packed) & 3u) == 0);
    if (vec) {
        const int64_t total4 = (rows * k) >> 2;
        const unsigned grid = tile_grid((total4 + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4), false);
#define SSQ_EXPORT(SB) (alpha ? export_vec_kernel<SB, true><<<grid, SSQ_THREADS, 0, st>>>(a) \
                              : export_vec_kernel<SB, false><<<grid, SSQ_THREADS, 0, st>>>(a))
        switch (sbits) { case 1: SSQ_EXPORT(1); break; case 2: SSQ_EXPORT(2); break; case 4: SSQ_EXPORT(4); break; default: SSQ_EXPORT(8); }
#undef SSQ_EXPORT
    } else {
        const int64_t total = rows * a.row_bytes;
        const int grid = grid_for((total + SSQ_THREADS - 1) / SSQ_THREADS);
        if (alpha) export_byte_kernel<true><<<grid, SSQ_THREADS, 0, st>>>(a, sbits);
        else export_byte_kernel<false><<<grid, SSQ_THREADS, 0, st>>>(a, sbits);
    }
    return launch_status();
}

extern "C" int ssq_import_codes(const uint8_t* packed, const float* in_scale, const float* delta, const float* zero_point,
                                float* w_q, int64_t rows, int64_t k, int64_t inner, int64_t nchan, float qmin,
                                int n_bits, void* stream) {
    if (rows == 0 || k == 0) return SSQ_OK;
    if (!packed || !delta || !zero_point || !w_q) return SSQ_ERR_NULL;
    if (n_bits < 1 || n_bits > 8 || !layout_ok(rows, k, inner, nchan)) return SSQ_ERR_SIZE;
    const int sbits = storage_bits(n_bits);
    ImportArgs a{packed, in_scale, delta, zero_point, w_q, rows, k, inner, nchan, ssq_packed_row_bytes(k, n_bits), qmin};
    cudaStream_t st = (cudaStream_t)stream;
    const int E = 32 / sbits;
    const bool vec = (k % E == 0) && (inner % 4 == 0) && !in_scale && aligned16(w_q) && ((reinterpret_cast<uintptr_t>(packed) & 3u) == 0);
    if (vec) {
        const int64_t total4 = (rows * k) >> 2;
        const unsigned grid = tile_grid((total4 + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4), false);
        switch (sbits) {
            case 1: import_vec_kernel<1><<<grid, SSQ_THREADS, 0, st>>>(a); break;
            case 2: import_vec_kernel<2><<<grid, SSQ_THREADS, 0, st>>>(a); break;
            case 4: import_vec_kernel<4><<<grid, SSQ_THREADS, 0, st>>>(a); break;
            default: import_vec_kernel<8><<<grid, SSQ_THREADS, 0, st>>>(a);
        }
    } else {
        const int grid = grid_for((rows * k + SSQ_THREADS * 4 - 1) / (SSQ_THREADS * 4));
        import_elem_kernel<<<grid, SSQ_THREADS, 0, st>>>(a, sbits);
    }
    return launch_status();
}
