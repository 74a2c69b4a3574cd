// mn_post.cuh -- the step after the merge path, on the device:
//   * bilinear resize of the class / sameness maps to the segmentation size
//     (egs/cityscape/local/segment.py:116-123: cv2.resize(maps, seg_size), INTER_LINEAR);
//   * nearest-neighbour resize of the instance masks back to the image size
//     (egs/cityscape/local/segment.py:147-149: cv2.resize(mask, (w, h), interpolation=cv2.INTER_NEAREST));
//   * COCO run-length encoding of every instance of a label mask
//     (egs/cityscape/local/segment.py:165-186, egs/coco/local/segment.py:190-204:
//      maskUtils.encode(np.asfortranarray(mask == i)) for i = 1..n).
// pycocotools is a third-party dependency that is neither vendored under /root/reference nor pinned by
// its requirements.txt; the encoding below restates the published cocoapi common/maskApi.c
// (rleEncode: column-major runs starting with a zero run; rleToString: LEB128-like, 5 payload bits per
// ASCII character from 48, counts from the fourth on delta-coded against the count two back).
// All integer work: results are compared bit for bit with the CPU restatement in oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>
#include <vector>

// ---- nearest-neighbour resize ------------------------------------------------------------------
// OpenCV resizeNN: ifx = 1 / (dst / (double)src); sx = min(cvFloor(x * ifx), src - 1); one IEEE double
// product and a floor per index -- evaluated on the device exactly as on the host
__global__ void mn_resize_nearest_kernel(const int* __restrict__ in, int* __restrict__ out, int H, int W, int OH, int OW,
                                         double ifx, double ify) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  int sy = (int)floor(__dmul_rn((double)y, ify));
  sy = sy < H - 1 ? sy : H - 1;
  const int* row = in + ((size_t)b * H + sy) * W;
  int* orow = out + ((size_t)b * OH + y) * OW;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < OW; x += gridDim.x * blockDim.x) {
    int sx = (int)floor(__dmul_rn((double)x, ifx));
    sx = sx < W - 1 ? sx : W - 1;
    orow[x] = row[sx];
  }
}

static cudaError_t mn_resize_nearest_launch(const int* d_in, int B, int H, int W, int* d_out, int OH, int OW, cudaStream_t s) {
  const double ifx = 1.0 / ((double)OW / (double)W), ify = 1.0 / ((double)OH / (double)H);
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    if (OH > 65535) return cudaErrorInvalidValue;
    dim3 g((unsigned)((OW + 255) / 256 < 64 ? (OW + 255) / 256 : 64), (unsigned)OH, (unsigned)nb);
    mn_resize_nearest_kernel<<<g, 256, 0, s>>>(d_in + (size_t)b0 * H * W, d_out + (size_t)b0 * OH * OW, H, W, OH, OW, ifx, ify);
  }
  return cudaGetLastError();
}

// ---- bilinear resize of the probability maps ---------------------------------------------------------
// cv2.resize(maps, (out_w, out_h)) with the default INTER_LINEAR on float32 maps of 2 or >= 5 channels
// (egs/cityscape/local/segment.py:116-123 resizes the C- and K-channel maps to the segmentation size), restated from
// OpenCV's resize.cpp: per destination column  fx = (float)((dx + 0.5) * (src_w / (double)dst_w) - 0.5), sx = floor(fx),
// fx -= sx (float), and (fx, sx) = (0, 0) left of the image, (0, src_w - 1) at or beyond its last column; per
// destination row the same WITHOUT that clamp of the weight -- the two source rows are clipped to the image instead;
// value = (S[sy0][sx] * (1 - fx) + S[sy0][sx + 1] * fx) * (1 - fy) + (S[sy1][sx] * (1 - fx) + S[sy1][sx + 1] * fx) * fy,
// every product and sum rounded to float on its own (no FMA: the generic many-channel path of the OpenCV in this image
// evaluates it that way; its 1-, 3- and 4-channel paths round differently and are not claimed).  Bit-identical to cv2
// 4.13 on the test matrix (tests/test_post.py).  Planar layout: d_in [planes][H][W] -> d_out [planes][OH][OW].
__global__ void __launch_bounds__(256) mn_resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                                 int OH, int OW, double scale_x, double scale_y) {
  const int y = blockIdx.y;
  const size_t pl = blockIdx.z;
  float fy = (float)(((double)y + 0.5) * scale_y - 0.5);
  const int sy = (int)floorf(fy);
  fy = __fsub_rn(fy, (float)sy);
  const int y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
  const float b0 = __fsub_rn(1.0f, fy), b1 = fy;
  const float* r0 = in + (pl * H + y0) * W;
  const float* r1 = in + (pl * H + y1) * W;
  float* orow = out + (pl * OH + y) * OW;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < OW; x += gridDim.x * blockDim.x) {
    float fx = (float)(((double)x + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx >= W - 1) { fx = 0.0f; sx = W - 1; }
    const int x1 = min(sx + 1, W - 1);
    const float a0 = __fsub_rn(1.0f, fx), a1 = fx;
    const float h0 = __fadd_rn(__fmul_rn(r0[sx], a0), __fmul_rn(r0[x1], a1));
    const float h1 = __fadd_rn(__fmul_rn(r1[sx], a0), __fmul_rn(r1[x1], a1));
    orow[x] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
  }
}
static cudaError_t mn_resize_bilinear_launch(const float* d_in, long long planes, int H, int W, float* d_out, int OH, int OW, cudaStream_t s) {
  if (OH > 65535) return cudaErrorInvalidValue;
  const double sx = (double)W / (double)OW, sy = (double)H / (double)OH;
  for (long long p0 = 0; p0 < planes; p0 += 65535) {
    const int np = (int)(planes - p0 < 65535 ? planes - p0 : 65535);
    dim3 g((unsigned)((OW + 255) / 256 < 64 ? (OW + 255) / 256 : 64), (unsigned)OH, (unsigned)np);
    mn_resize_bilinear_kernel<<<g, 256, 0, s>>>(d_in + (size_t)p0 * H * W, d_out + (size_t)p0 * OH * OW, H, W, OH, OW, sx, sy);
  }
  return cudaGetLastError();
}

// ---- COCO RLE ------------------------------------------------------------------------------------
// 1. column-major copy of the label mask (32x32 tiles through shared memory: both sides coalesced)
__global__ void mn_rle_transpose_kernel(const int* __restrict__ in, int* __restrict__ out, int H, int W, int n) {
  __shared__ int tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < H && c < W) {
      const int v = in[(size_t)r * W + c];
      tile[i][threadIdx.x] = (v < 0 || v > n) ? 0 : v;  // (labels outside 0..n belong to no instance)
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < H && c < W) out[(size_t)c * H + r] = tile[threadIdx.x][i];
  }
}
// 2. run boundaries of the non-zero labels in column-major order: low word = "a run starts here",
//    high word = "a run ends here" (one 64-bit scan numbers both; starts and ends alternate, so the
//    k-th end closes the k-th start)
__global__ void mn_rle_flags_kernel(const int* __restrict__ lt, unsigned long long* __restrict__ flags, long long a) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a; j += (long long)gridDim.x * blockDim.x) {
    const int v = lt[j];
    unsigned long long f = 0;
    if (v != 0) {
      if (j == 0 || lt[j - 1] != v) f |= 1ull;
      if (j == a - 1 || lt[j + 1] != v) f |= 1ull << 32;
    }
    flags[j] = f;
  }
}
__global__ void mn_rle_runs_kernel(const int* __restrict__ lt, const unsigned long long* __restrict__ flags,
                                   const unsigned long long* __restrict__ scan, int* __restrict__ run_label,
                                   int* __restrict__ run_start, int* __restrict__ run_end, int* __restrict__ run_idx, long long a) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a; j += (long long)gridDim.x * blockDim.x) {
    const unsigned long long f = flags[j], sc = scan[j];
    if (f & 1ull) {
      const unsigned k = (unsigned)(sc & 0xffffffffull);
      run_label[k] = lt[j]; run_start[k] = (int)j; run_idx[k] = (int)k;
    }
    if (f >> 32) run_end[(unsigned)(sc >> 32)] = (int)(j + 1);
  }
}
// 3. (runs sorted by label, stable) per instance v = 1..n: first sorted position with label >= v
__global__ void mn_rle_segments_kernel(const int* __restrict__ sorted_label, int R, int n, int* __restrict__ seg_begin /* n + 2 */) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= n + 1; v += gridDim.x * blockDim.x) {
    int lo = 0, hi = R;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (sorted_label[mid] < v) lo = mid + 1; else hi = mid; }
    seg_begin[v] = lo;
  }
}
// Count slots of instance v start at 2 * seg_begin[v] + (v - 1): two per run plus one for the trailing
// zero run (unused when the last run touches the end of the mask).
__device__ __forceinline__ int mn_rle_nchars(long long x) {
  int nc = 0;
  bool more = true;
  while (more) {
    const int c = (int)(x & 0x1f);
    x >>= 5;
    more = (c & 0x10) ? x != -1 : x != 0;
    nc++;
  }
  return nc;
}
__global__ void mn_rle_counts_kernel(const int* __restrict__ sorted_label, const int* __restrict__ sorted_run,
                                     const int* __restrict__ run_start, const int* __restrict__ run_end,
                                     const int* __restrict__ seg_begin, int R, int n, long long a,
                                     unsigned* __restrict__ cnts, int* __restrict__ used) {
  // one thread per sorted run: its zero gap and its length; the last run of an instance also decides the trailing slot
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < R + n; q += gridDim.x * blockDim.x) {
    if (q < R) {
      const int v = sorted_label[q], k = sorted_run[q];
      const int first = seg_begin[v], t = q - first;
      const int s = run_start[k], e = run_end[k];
      const int prev_e = t > 0 ? run_end[sorted_run[q - 1]] : 0;
      const size_t base = 2 * (size_t)first + (size_t)(v - 1);
      cnts[base + 2 * t] = (unsigned)(s - prev_e); used[base + 2 * t] = 1;
      cnts[base + 2 * t + 1] = (unsigned)(e - s); used[base + 2 * t + 1] = 1;
      if (q + 1 == seg_begin[v + 1]) {  // last run of the instance
        const size_t tr = base + 2 * (size_t)(t + 1);
        cnts[tr] = (unsigned)(a - e); used[tr] = (a - e) > 0 ? 1 : 0;
      }
    } else {  // an instance without pixels (possible after a resize): the single count a
      const int v = q - R + 1;
      if (seg_begin[v] == seg_begin[v + 1]) {
        const size_t base = 2 * (size_t)seg_begin[v] + (size_t)(v - 1);
        cnts[base] = (unsigned)a; used[base] = 1;
      }
    }
  }
}
__global__ void mn_rle_nchars_kernel(const unsigned* __restrict__ cnts, const int* __restrict__ used, const int* __restrict__ seg_begin,
                                     const int* __restrict__ slot_inst, long long nslots, long long* __restrict__ nchar) {
  for (long long sl = (long long)blockIdx.x * blockDim.x + threadIdx.x; sl < nslots; sl += (long long)gridDim.x * blockDim.x) {
    int nc = 0;
    if (used[sl]) {
      const int v = slot_inst[sl];
      const long long i = sl - (2 * (long long)seg_begin[v] + (v - 1));
      long long x = (long long)cnts[sl];
      if (i > 2) x -= (long long)cnts[sl - 2];
      nc = mn_rle_nchars(x);
    }
    nchar[sl] = nc;
  }
}
// instance of every slot (slots of v: [2 seg_begin[v] + v - 1, 2 seg_begin[v+1] + v))
__global__ void mn_rle_slot_inst_kernel(const int* __restrict__ seg_begin, int n, int* __restrict__ slot_inst) {
  const int v = blockIdx.x + 1;
  if (v > n) return;
  const long long b = 2 * (long long)seg_begin[v] + (v - 1), e = 2 * (long long)seg_begin[v + 1] + v;
  for (long long sl = b + threadIdx.x; sl < e; sl += blockDim.x) slot_inst[sl] = v;
}
__global__ void mn_rle_write_kernel(const unsigned* __restrict__ cnts, const int* __restrict__ used, const int* __restrict__ seg_begin,
                                    const int* __restrict__ slot_inst, const long long* __restrict__ char_ofs, long long nslots,
                                    unsigned char* __restrict__ out, long long cap) {
  for (long long sl = (long long)blockIdx.x * blockDim.x + threadIdx.x; sl < nslots; sl += (long long)gridDim.x * blockDim.x) {
    if (!used[sl]) continue;
    const int v = slot_inst[sl];
    const long long i = sl - (2 * (long long)seg_begin[v] + (v - 1));
    long long x = (long long)cnts[sl];
    if (i > 2) x -= (long long)cnts[sl - 2];
    long long p = char_ofs[sl];
    bool more = true;
    while (more) {
      int c = (int)(x & 0x1f);
      x >>= 5;
      more = (c & 0x10) ? x != -1 : x != 0;
      if (more) c |= 0x20;
      if (p < cap) out[p] = (unsigned char)(c + 48);
      p++;
    }
  }
}
__global__ void mn_rle_offsets_kernel(const int* __restrict__ seg_begin, const long long* __restrict__ char_ofs, int n, long long nslots,
                                      long long total, long long* __restrict__ offsets) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x + 1; v <= n + 1; v += gridDim.x * blockDim.x) {
    const long long sl = 2 * (long long)seg_begin[v] + (v - 1);
    offsets[v - 1] = sl < nslots ? char_ofs[sl] : total;
  }
}

// the whole encoding of one mask; d_mask on the device.  Returns the number of bytes the strings need in *total.
static cudaError_t mn_coco_rle_device(const int* d_mask, int H, int W, int n, unsigned char* d_counts, long long cap,
                                      long long* d_offsets /* n + 1 */, long long* total, cudaStream_t s) {
  const long long a = (long long)H * W;
  cudaError_t e;
  int *lt = nullptr, *run_label = nullptr, *run_start = nullptr, *run_end = nullptr, *run_idx = nullptr;
  int *sorted_label = nullptr, *sorted_run = nullptr, *seg_begin = nullptr, *used = nullptr, *slot_inst = nullptr;
  long long* nchar = nullptr;
  unsigned long long *flags = nullptr, *scan = nullptr;
  unsigned* cnts = nullptr;
  long long* char_ofs = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(lt); cudaFree(run_label); cudaFree(run_start); cudaFree(run_end); cudaFree(run_idx); cudaFree(sorted_label);
    cudaFree(sorted_run); cudaFree(seg_begin); cudaFree(used); cudaFree(slot_inst); cudaFree(nchar); cudaFree(flags);
    cudaFree(scan); cudaFree(cnts); cudaFree(char_ofs); cudaFree(tmp);
  };
#define MN_POST_OK(x) do { e = (x); if (e != cudaSuccess) { cleanup(); return e; } } while (0)
  MN_POST_OK(cudaMalloc(&lt, a * 4));
  MN_POST_OK(cudaMalloc(&flags, a * 8));
  MN_POST_OK(cudaMalloc(&scan, a * 8));
  {
    dim3 g((unsigned)((W + 31) / 32), (unsigned)((H + 31) / 32)), b(32, 8);
    mn_rle_transpose_kernel<<<g, b, 0, s>>>(d_mask, lt, H, W, n);
  }
  const int gs = (int)std::min<long long>(148 * 8, (a + 255) / 256);
  mn_rle_flags_kernel<<<gs, 256, 0, s>>>(lt, flags, a);
  size_t tb = 0;
  MN_POST_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, flags, scan, (int)a, s));
  MN_POST_OK(cudaMalloc(&tmp, tb));
  MN_POST_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, flags, scan, (int)a, s));
  unsigned long long last_scan = 0, last_flag = 0;
  MN_POST_OK(cudaMemcpyAsync(&last_scan, scan + (a - 1), 8, cudaMemcpyDeviceToHost, s));
  MN_POST_OK(cudaMemcpyAsync(&last_flag, flags + (a - 1), 8, cudaMemcpyDeviceToHost, s));
  MN_POST_OK(cudaStreamSynchronize(s));
  const int R = (int)((last_scan + last_flag) & 0xffffffffull);
  const size_t Ra = (size_t)(R > 0 ? R : 1);
  MN_POST_OK(cudaMalloc(&run_label, Ra * 4)); MN_POST_OK(cudaMalloc(&run_start, Ra * 4));
  MN_POST_OK(cudaMalloc(&run_end, Ra * 4)); MN_POST_OK(cudaMalloc(&run_idx, Ra * 4));
  MN_POST_OK(cudaMalloc(&sorted_label, Ra * 4)); MN_POST_OK(cudaMalloc(&sorted_run, Ra * 4));
  MN_POST_OK(cudaMalloc(&seg_begin, ((size_t)n + 2) * 4));
  mn_rle_runs_kernel<<<gs, 256, 0, s>>>(lt, flags, scan, run_label, run_start, run_end, run_idx, a);
  if (R > 0) {
    int bits = 1;
    while ((1ll << bits) <= (long long)n + 1 && bits < 31) bits++;
    size_t tb2 = 0;
    MN_POST_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb2, run_label, sorted_label, run_idx, sorted_run, R, 0, bits, s));
    if (tb2 > tb) { cudaFree(tmp); tmp = nullptr; MN_POST_OK(cudaMalloc(&tmp, tb2)); tb = tb2; }
    MN_POST_OK(cub::DeviceRadixSort::SortPairs(tmp, tb2, run_label, sorted_label, run_idx, sorted_run, R, 0, bits, s));
  }
  mn_rle_segments_kernel<<<(n + 2 + 255) / 256, 256, 0, s>>>(sorted_label, R, n, seg_begin);
  const long long nslots = 2 * (long long)R + n;
  const size_t ns = (size_t)(nslots > 0 ? nslots : 1);
  MN_POST_OK(cudaMalloc(&cnts, ns * 4)); MN_POST_OK(cudaMalloc(&used, ns * 4)); MN_POST_OK(cudaMalloc(&slot_inst, ns * 4));
  MN_POST_OK(cudaMalloc(&nchar, ns * 8)); MN_POST_OK(cudaMalloc(&char_ofs, ns * 8));
  MN_POST_OK(cudaMemsetAsync(used, 0, ns * 4, s));
  *total = 0;
  if (n > 0) {
    mn_rle_slot_inst_kernel<<<n, 128, 0, s>>>(seg_begin, n, slot_inst);
    const int gq = (int)std::min<long long>(148 * 8, ((long long)R + n + 255) / 256);
    mn_rle_counts_kernel<<<gq, 256, 0, s>>>(sorted_label, sorted_run, run_start, run_end, seg_begin, R, n, a, cnts, used);
    const int gsl = (int)std::min<long long>(148 * 8, (nslots + 255) / 256);
    mn_rle_nchars_kernel<<<gsl, 256, 0, s>>>(cnts, used, seg_begin, slot_inst, nslots, nchar);
    size_t tb3 = 0;
    MN_POST_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb3, nchar, char_ofs, (int)nslots, s));
    if (tb3 > tb) { cudaFree(tmp); tmp = nullptr; MN_POST_OK(cudaMalloc(&tmp, tb3)); tb = tb3; }
    MN_POST_OK(cub::DeviceScan::ExclusiveSum(tmp, tb3, nchar, char_ofs, (int)nslots, s));
    long long lo = 0, ln = 0;
    MN_POST_OK(cudaMemcpyAsync(&lo, char_ofs + (nslots - 1), 8, cudaMemcpyDeviceToHost, s));
    MN_POST_OK(cudaMemcpyAsync(&ln, nchar + (nslots - 1), 8, cudaMemcpyDeviceToHost, s));
    MN_POST_OK(cudaStreamSynchronize(s));
    *total = lo + ln;
    mn_rle_write_kernel<<<gsl, 256, 0, s>>>(cnts, used, seg_begin, slot_inst, char_ofs, nslots, d_counts, cap);
    mn_rle_offsets_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(seg_begin, char_ofs, n, nslots, *total, d_offsets);
  } else {
    MN_POST_OK(cudaMemsetAsync(d_offsets, 0, 8, s));
  }
  MN_POST_OK(cudaStreamSynchronize(s));
  e = cudaGetLastError();
  cleanup();
#undef MN_POST_OK
  return e;
}
