// vi_ingest.cuh -- frame ingest on the device: the two conversions the reference does on the host before
// the hot path sees a mono uint8 frame.  Both are pure HBM streaming: 128-bit coalesced loads with
// many bytes in flight per thread, 128-bit stores, grid sized in multiples of the SM count.
//
//   ARGB32 -> gray  segmentation.qimage_to_gray_array (segmentation.py:15-23): QImage ARGB32 is B,G,R,A in
//                   memory; the reference reverses the first three channels and hands them to
//                   cv2.cvtColor(BGR2GRAY), so OpenCV's B weight meets R and its R weight meets B:
//                   gray = (R*3735 + G*19235 + B*9798 + 16384) >> 15  (15-bit fixed point).
//   uint16 -> uint8 ImageWidget.load_image (indexing_ui.py:153-155): (arr / 256).astype(uint8) == v >> 8.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vi {

constexpr int kIngestThreads = 256;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ unsigned gray_of_bgra(unsigned px) {
    const unsigned b = px & 255u, g = (px >> 8) & 255u, r = (px >> 16) & 255u;
    return (r * 3735u + g * 19235u + b * 9798u + 16384u) >> 15;
}

__device__ __forceinline__ unsigned gray4(const uint4& v) {
    return gray_of_bgra(v.x) | (gray_of_bgra(v.y) << 8) | (gray_of_bgra(v.z) << 16) | (gray_of_bgra(v.w) << 24);
}

// One item = 16 output pixels = 64 source bytes.  Rows are processed as whole items when pointers and pitches
// are 16-byte aligned and W is a multiple of 16 (the fast path: frames are 4096 wide); anything else takes the
// scalar kernel.  n_items = n_images * H * (W / 16).
__global__ void __launch_bounds__(kIngestThreads) ingest_argb32_v16(const uint8_t* __restrict__ src, long long src_pitch,
                                                                     long long src_stride, uint8_t* __restrict__ dst,
                                                                     long long dst_pitch, long long dst_stride, int W, int H,
                                                                     long long n_items) {
    const int ipr = W >> 4;                      // items per row
    const long long step = (long long)gridDim.x * kIngestThreads;
    for (long long it0 = (long long)blockIdx.x * kIngestThreads + threadIdx.x; it0 < n_items; it0 += 2 * step) {
        // two items per pass: eight 16-byte loads in flight per thread
        const long long it1 = it0 + step;
        const bool two = it1 < n_items;
        const long long r0 = it0 / ipr, r1 = two ? it1 / ipr : r0;
        const int c0 = (int)(it0 - r0 * ipr), c1 = two ? (int)(it1 - r1 * ipr) : c0;
        const long long i0 = r0 / H, i1 = r1 / H;
        const int y0 = (int)(r0 - i0 * H), y1 = (int)(r1 - i1 * H);
        const uint4* s0 = reinterpret_cast<const uint4*>(src + i0 * src_stride + (long long)y0 * src_pitch) + 4 * c0;
        const uint4* s1 = reinterpret_cast<const uint4*>(src + i1 * src_stride + (long long)y1 * src_pitch) + 4 * c1;
        const uint4 a0 = ldg_stream(s0), a1 = ldg_stream(s0 + 1), a2 = ldg_stream(s0 + 2), a3 = ldg_stream(s0 + 3);
        uint4 b0 = a0, b1 = a1, b2 = a2, b3 = a3;
        if (two) { b0 = ldg_stream(s1); b1 = ldg_stream(s1 + 1); b2 = ldg_stream(s1 + 2); b3 = ldg_stream(s1 + 3); }
        stg_stream(reinterpret_cast<uint4*>(dst + i0 * dst_stride + (long long)y0 * dst_pitch) + c0,
                   make_uint4(gray4(a0), gray4(a1), gray4(a2), gray4(a3)));
        if (two)
            stg_stream(reinterpret_cast<uint4*>(dst + i1 * dst_stride + (long long)y1 * dst_pitch) + c1,
                       make_uint4(gray4(b0), gray4(b1), gray4(b2), gray4(b3)));
    }
}

__global__ void __launch_bounds__(kIngestThreads) ingest_argb32_scalar(const uint8_t* __restrict__ src, long long src_pitch,
                                                                        long long src_stride, uint8_t* __restrict__ dst,
                                                                        long long dst_pitch, long long dst_stride, int W, int H,
                                                                        long long n_px) {
    for (long long e = (long long)blockIdx.x * kIngestThreads + threadIdx.x; e < n_px; e += (long long)gridDim.x * kIngestThreads) {
        const long long r = e / W;
        const int x = (int)(e - r * W);
        const long long i = r / H;
        const int y = (int)(r - i * H);
        const uint8_t* p = src + i * src_stride + (long long)y * src_pitch + 4ll * x;
        const unsigned px = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
        dst[i * dst_stride + (long long)y * dst_pitch + x] = (uint8_t)gray_of_bgra(px);
    }
}

// uint16 -> uint8: one item = 16 pixels = 32 source bytes; high bytes picked with two byte permutes per word pair.
__device__ __forceinline__ unsigned hi4(unsigned a, unsigned b) { return __byte_perm(a, b, 0x7531); }   // high bytes of 4 uint16

__global__ void __launch_bounds__(kIngestThreads) ingest_gray16_v16(const uint8_t* __restrict__ src, long long src_pitch,
                                                                     long long src_stride, uint8_t* __restrict__ dst,
                                                                     long long dst_pitch, long long dst_stride, int W, int H,
                                                                     long long n_items) {
    const int ipr = W >> 4;
    const long long step = (long long)gridDim.x * kIngestThreads;
    for (long long it0 = (long long)blockIdx.x * kIngestThreads + threadIdx.x; it0 < n_items; it0 += 4 * step) {
        uint4 a[4][2];
        long long dsto[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long it = it0 + k * step;
            ok[k] = it < n_items;
            const long long itc = ok[k] ? it : it0;
            const long long r = itc / ipr;
            const int c = (int)(itc - r * ipr);
            const long long i = r / H;
            const int y = (int)(r - i * H);
            const uint4* s = reinterpret_cast<const uint4*>(src + i * src_stride + (long long)y * src_pitch) + 2 * c;
            a[k][0] = ldg_stream(s); a[k][1] = ldg_stream(s + 1);
            dsto[k] = i * dst_stride + (long long)y * dst_pitch + 16ll * c;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (ok[k])
                stg_stream(reinterpret_cast<uint4*>(dst + dsto[k]),
                           make_uint4(hi4(a[k][0].x, a[k][0].y), hi4(a[k][0].z, a[k][0].w), hi4(a[k][1].x, a[k][1].y), hi4(a[k][1].z, a[k][1].w)));
    }
}

__global__ void __launch_bounds__(kIngestThreads) ingest_gray16_scalar(const uint8_t* __restrict__ src, long long src_pitch,
                                                                        long long src_stride, uint8_t* __restrict__ dst,
                                                                        long long dst_pitch, long long dst_stride, int W, int H,
                                                                        long long n_px) {
    for (long long e = (long long)blockIdx.x * kIngestThreads + threadIdx.x; e < n_px; e += (long long)gridDim.x * kIngestThreads) {
        const long long r = e / W;
        const int x = (int)(e - r * W);
        const long long i = r / H;
        const int y = (int)(r - i * H);
        const uint8_t* p = src + i * src_stride + (long long)y * src_pitch + 2ll * x;
        dst[i * dst_stride + (long long)y * dst_pitch + x] = p[1];          // little-endian uint16: the high byte
    }
}

}  // namespace vi
