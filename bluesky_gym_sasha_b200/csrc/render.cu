// render.cu -- rgb_array frames of ONE env: a primitive list rasterised on the device.
//
// The reference draws its frames with pygame (`_render_frame`, e.g. horizontal_cr_env.py:277-395: lines of a given width,
// circles / rings, a filled rectangle, filled or outlined polygons on a 512 x 512 canvas) and only ever shows them in a
// window (`render_mode="human"`); `rgb_array` is declared in the metadata but not implemented there.  Here the host builds
// the same draw calls from the env's state (bluesky_gym_sasha_b200/render.py mirrors each env's _render_frame) and this
// kernel paints them: one thread per pixel, primitives staged in shared memory, painter's order (a later primitive
// overwrites an earlier one), pixel centres at (x + 0.5, y + 0.5).  Coverage rules (stated, not pygame's: pygame is not in
// the image): a LINE covers pixels within max(width, 1) / 2 of the segment; a RING of radius r and width w covers
// r - w <= d <= r, width 0 fills the disc; a RECT is filled; a POLYGON is a run of EDGE records closed by an EDGE_END
// record and is filled by the even-odd rule.  oracle/raster.py restates the rules in NumPy.
#include <cuda_runtime.h>
#include <stdint.h>

#include "bsg_internal.h"

namespace bsg {

constexpr int kMaxPrims = 1024;          // 32 KB of shared memory

__global__ void __launch_bounds__(256) render_kernel(const float* __restrict__ prims, const int n_prims, const int width,
                                                     const int height, const uint32_t bg, uint8_t* __restrict__ rgb) {
    __shared__ float s_p[kMaxPrims * BSG_PRIM_FLOATS];
    for (int k = threadIdx.x; k < n_prims * BSG_PRIM_FLOATS; k += blockDim.x) s_p[k] = prims[k];
    __syncthreads();
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= width || py >= height) return;
    const float x = (float)px + 0.5f, y = (float)py + 0.5f;
    uint32_t col = bg;
    int parity = 0;
    for (int k = 0; k < n_prims; ++k) {
        const float* p = s_p + k * BSG_PRIM_FLOATS;
        const int type = (int)p[0];
        const float x0 = p[1], y0 = p[2], x1 = p[3], y1 = p[4], w = p[5];
        const uint32_t c = __float_as_uint(p[6]);
        bool hit = false;
        if (type == BSG_PRIM_LINE) {
            const float dx = x1 - x0, dy = y1 - y0, l2 = dx * dx + dy * dy;
            float t = l2 > 0.0f ? ((x - x0) * dx + (y - y0) * dy) / l2 : 0.0f;
            t = fminf(1.0f, fmaxf(0.0f, t));
            const float ex = x - (x0 + t * dx), ey = y - (y0 + t * dy), hw = 0.5f * fmaxf(w, 1.0f);
            hit = ex * ex + ey * ey <= hw * hw;
        } else if (type == BSG_PRIM_RING) {
            const float ex = x - x0, ey = y - y0, d2 = ex * ex + ey * ey, r = x1, ri = fmaxf(r - w, 0.0f);
            hit = d2 <= r * r && (w <= 0.0f || d2 >= ri * ri);
        } else if (type == BSG_PRIM_RECT) {
            hit = x >= x0 && x < x1 && y >= y0 && y < y1;
        } else if (type == BSG_PRIM_EDGE || type == BSG_PRIM_EDGE_END) {
            if ((y0 > y) != (y1 > y) && x < (x1 - x0) * (y - y0) / (y1 - y0) + x0) parity ^= 1;
            if (type == BSG_PRIM_EDGE_END) { hit = parity != 0; parity = 0; }
        }
        if (hit) col = c;
    }
    uint8_t* o = rgb + 3 * ((size_t)py * width + px);
    o[0] = (uint8_t)(col & 0xffu); o[1] = (uint8_t)((col >> 8) & 0xffu); o[2] = (uint8_t)((col >> 16) & 0xffu);
}

}  // namespace bsg

extern "C" int bsg_render(const float* d_prims, int32_t n_prims, int32_t width, int32_t height, uint32_t background_rgb,
                          uint8_t* d_rgb, void* stream) {
    if (n_prims < 0 || n_prims > bsg::kMaxPrims) return bsg_fail(BSG_EINVAL, "bsg_render: n_prims must be in [0, 1024]");
    if (width <= 0 || height <= 0 || width > 8192 || height > 8192) return bsg_fail(BSG_EINVAL, "bsg_render: bad frame size");
    if ((n_prims > 0 && !d_prims) || !d_rgb) return bsg_fail(BSG_EINVAL, "bsg_render: null pointer");
    dim3 grid((width + 15) / 16, (height + 15) / 16);
    bsg::render_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_prims, n_prims, width, height, background_rgb, d_rgb);
    return bsg_cuda_check(cudaGetLastError(), "bsg_render launch");
}
