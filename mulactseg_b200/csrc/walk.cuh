// Balanced walk over the strip rows of a batch of images (shared by the streaming kernels).
//
// A STRIP ROW is 32*VEC consecutive pixels of one image row; lane l of a warp owns VEC consecutive
// columns.  Strip rows are linearised (image, strip, y) and cut into one contiguous range per warp of a
// single-wave persistent grid, so that every warp gets the same number of rows (+-1) whatever the batch
// size and walks DOWN its strip (staying inside one superpixel for many rows).
#pragma once

namespace mas {

// position of a strip row in the (image, strip, y) order
struct Cursor {
    int img, strip, y;
    __device__ __forceinline__ void seek(long long row, int strips, int H) {
        const long long per_img = (long long)strips * H;
        img = (int)(row / per_img);
        const long long rem = row - (long long)img * per_img;
        strip = (int)(rem / H);
        y = (int)(rem - (long long)strip * H);
    }
    // returns 0 = same strip, 1 = next strip of the image, 2 = next image
    __device__ __forceinline__ int advance(int strips, int H) {
        if (++y < H) return 0;
        y = 0;
        if (++strip < strips) return 1;
        strip = 0;
        ++img;
        return 2;
    }
};

// rows [r0, r1) of warp `warp` out of `n_warps`: consecutive, cover [0,total), sizes differ by at most one
__device__ __forceinline__ void warp_range(long long total_rows, long long warp, long long n_warps, long long& r0, long long& r1) {
    r0 = (long long)(((unsigned long long)total_rows * (unsigned long long)warp) / (unsigned long long)n_warps);
    r1 = (long long)(((unsigned long long)total_rows * (unsigned long long)(warp + 1)) / (unsigned long long)n_warps);
}

}  // namespace mas
