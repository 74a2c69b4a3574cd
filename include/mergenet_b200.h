/*
 * mergenet_b200.h -- C ABI of the B200-native MergeNet merge segmenter (libmergenet_b200.so).
 *
 * Plain C, plain pointers and sizes, no torch types.  The library is CUDA-only (sm_100a): every
 * entry point that computes fails with MN_ERR_CUDA when no device is usable; there is no CPU path.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   c_run_segmentation      <- utils/csegment/segment.cc:742-765 (extern "C", identical signature);
 *                              it is what utils/csegment/c_segment.pyx:16-25,69-78 binds.
 *   mn_segment_batch_*      <- additive: the same operation over B images of one shape (device or
 *                              host buffers), used by the Python facade and the benchmark.
 */
#ifndef MERGENET_B200_H
#define MERGENET_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes (0 = ok). */
enum {
  MN_STATUS_OK = 0,
  MN_STATUS_BAD_ARG = 1,
  MN_STATUS_PL_POOL = 2,
  MN_STATUS_Q_POOL = 3,
  MN_STATUS_TREE_POOL = 4,
  MN_STATUS_HASH_FULL = 5,
  MN_STATUS_INTERNAL = 6,
  MN_STATUS_CUDA = 7,
  MN_STATUS_LIMIT = 8,
  MN_STATUS_NO_BACKGROUND = 9 /* Mode B only: prune() found no class-0 object (utils/segmenter.py:351-375 raises) */
};

/*
 * Drop-in for the reference symbol (segment.cc:742-752).  NOTE width before height.
 *   class_pred   class_dim x H x W  fp32, C-contiguous, already clipped to [2^-23, 1-2^-23]
 *   adj_pred     offset_dim x H x W fp32; REWRITTEN IN PLACE when same_different_bias != 0
 *                (segment.cc:183-195)
 *   offset_list  offset_dim x 2 int32 (row delta, col delta)   (segment.cc:166-169)
 *   output       H x W int32, fully overwritten: 0 = class-0 objects, 1..n = instances
 *   object_class H*W int32: -1 everywhere, then the class of label k at [k-1] (segment.cc:497-509)
 * All buffers are HOST memory owned by the caller.  The work runs on the calling thread's CURRENT CUDA
 * device (cudaGetDevice), with one workspace cached per thread and reused while the shape stays the same
 * (freed when the thread exits, or by mn_shutdown()).  Returns nothing, like the reference; unlike the
 * reference it never calls exit() (segment.cc:40-43,666-673): on failure the outputs are left as (0, -1),
 * the code is readable through mn_last_error(), and -- because a caller that links this void symbol directly
 * (the reference's own c_segment.pyx) cannot see a status -- one line naming the status goes to stderr.
 * Nothing is printed on success (the reference's progress lines, segment.cc:540-569, are not reproduced).
 * Shape limits: those of mn_plan_create below.
 * With MN_TIE_ORDER=reference in the environment the call runs the tie-exact replay (mn_exact_segment_host below)
 * instead of the hot path: the reference's RAW arrays also on inputs whose partition depends on the order among
 * exactly equal priorities, at sequential speed -- a switch for callers that bind this symbol directly.
 */
void c_run_segmentation(float* class_pred, int class_dim, float* adj_pred, int offset_dim,
                        int img_width, int img_height, int num_classes, int* offset_list,
                        int* output, int* object_class, float same_different_bias,
                        float object_merge_factor, float merge_logprob_bias);

/* Frees what the calling thread cached for the host-buffer entries above and below (the drop-in's workspace,
 * the post-pass scratch).  Optional: thread exit does the same. */
void mn_shutdown(void);

/* Status of the last call on this thread, and a static description of a status code. */
int mn_last_error(void);
const char* mn_status_string(int status);

/* Number of usable CUDA devices (0 when the driver or a device is missing). */
int mn_device_count(void);

/* ---- batched interface ---------------------------------------------------------------------- */
#define MN_INPUT_CLIP 1
#define MN_INPUT_LOGITS 2

typedef struct mn_plan mn_plan; /* workspace for up to max_batch images of one shape on one GPU */

/* Per-image statistics of the last run (north star: round count, merges, per-round latency). */
typedef struct {
  int status;
  int fail_line;
  int n_instances;
  int n_init_entries;
  long long rounds, events, merges, restores, invalid_pops, solo_events;
  long long refills, flushes, splits, pairs, cuts_conflict, cuts_cascade, cuts_capacity;
  long long queue_chunks_used, pixel_pool_used, tree_nodes_used;
  long long requeues; /* guard entries replaced by an exact entry (lazy queue) */
  long long hash_overflow; /* records living in the hash overflow area */
  long long pixel_pool_collections; /* semi-space collections of the pixel-array pool */
  /* SM cycles of the image's CTA: total, and by phase (select, plan, accept, commit, hot-queue update,
   * flush, refill, split, solo merges, gc) */
  long long cycles_total;
  long long cycles[16]; /* ... + refill leaves / init / sort, select stage / classify / pixels */
} mn_image_stats;

/* Device time of the phases of the last batch (CUDA events on the plan's stream), milliseconds. */
typedef struct {
  float h2d_ms, edge_ms, record_init_sort_ms, merge_ms, label_ms, d2h_ms, total_ms, aggregate_ms;
  long long edge_launches, other_launches; /* kernels of this library launched by the last batch */
} mn_timings;

/* Bytes of device workspace one image of this shape needs (for sizing max_batch). */
size_t mn_workspace_bytes_per_image(int height, int width, int num_classes, int num_offsets);

/* Shape limits (MN_STATUS_BAD_ARG beyond them): height * width < 2^24 pixels, height * width * num_offsets <=
 * 2^25 record slots (record ids share a 32-bit hash word with a 6-bit fingerprint), num_classes < 256,
 * num_offsets <= 16.  Every BASELINE shape fits (1024 x 2048 x 10 = 21.0 M slots). */
int mn_plan_create(mn_plan** out, int max_batch, int height, int width, int num_classes,
                   int num_offsets, const int* offset_list /* K x 2 (drow, dcol) */, int device);
void mn_plan_destroy(mn_plan* plan);

/*
 * Segment `batch` images whose maps are ALREADY ON THE DEVICE of the plan.
 *   d_class [B][C][H][W] fp32, d_adj [B][K][H][W] fp32 (rewritten when same_different_bias != 0),
 *   d_mask [B][H][W] int32, d_object_class [B][H*W] int32, d_num_instances [B] int32.
 *   clip: input flags.  MN_INPUT_CLIP applies the wrapper's clip to [2^-23, 1-2^-23]
 *   (c_segment.pyx:53-55) on the fly; MN_INPUT_LOGITS says the maps are the network's raw outputs:
 *   the edge pass applies F.sigmoid (utils/inference_utils.py:43-44,95-96: 1 / (1 + exp(-x)) in fp32,
 *   as torch evaluates it on the device) and the clip while it reads them, so the probability maps
 *   never exist in memory.  0 = probabilities the caller has ALREADY clipped (a contract: values must lie in
 *   [2^-126, 1); only c_run_segmentation, which like the reference symbol takes any floats, checks the domain
 *   and falls back to the kernel that honours libm's special values for 0, 1, negatives and NaN).
 *   stream: a cudaStream_t (NULL = the plan's own stream).  The call returns after the work has
 *   completed (it synchronises the stream to read the per-image status words).
 */
int mn_segment_batch_device(mn_plan* plan, int batch, const float* d_class, float* d_adj,
                            int* d_mask, int* d_object_class, int* d_num_instances, int clip,
                            float same_different_bias, float object_merge_factor,
                            float merge_logprob_bias, void* stream);

/* Same with HOST buffers (pinned or pageable): the copies are part of the call. */
int mn_segment_batch_host(mn_plan* plan, int batch, const float* h_class, float* h_adj, int* h_mask,
                          int* h_object_class, int* h_num_instances, int clip,
                          float same_different_bias, float object_merge_factor,
                          float merge_logprob_bias);

int mn_plan_image_stats(mn_plan* plan, int image, mn_image_stats* out);
int mn_plan_timings(mn_plan* plan, mn_timings* out);
/* Total log-probability of the last run's segmentation of `image`, the quantity the reference prints
 * (segment.cc:314-350, ComputeTotalLogprobFromScratch; the maintained variant cc:272-287 prints the same number up
 * to float rounding) and does not return: out4 = {class term: sum over pixels of log p(class of the pixel's
 * instance; class 0 for background), sameness term: sum of log(s) over the in-image (pixel, offset) pairs inside one
 * instance, differentness term: sum of log(1 - s) over the pairs across two, class + object_merge_factor *
 * (differentness + sameness)}.  Evaluated on the GPU in float64 from the maps the run saw and its label mask (all
 * background pixels count as one region, as the mask shows them) by a streaming partition-statistics pass. */
int mn_plan_image_logprob(mn_plan* plan, int image, double* out4);

/* ---- Mode B: the semantics of the reference's pure-Python segmenter ------------------------------------------ */
/*
 * utils/segmenter.py::ObjectSegmenter.run_segmentation (py:432-483) for ONE image, host buffers: priority
 * (oml * omf + cdl + mlb) / (n1 * n2) (py:189-193), merge when the recomputed priority >= the popped one (py:470),
 * float64 class accumulators (py:51), heapq's own order among equal priorities, prune(prune_threshold) (py:351-375;
 * the reference always uses 200), labels in ascending surviving id, int64 mask (py:377-389).  The three inputs are
 * the LOGARITHMS the reference takes with NumPy -- np.log(class_probs) [C][H][W], np.log(same) and
 * np.log(1.0 - same) [K][H][W], float32 -- computed by the caller's NumPy so that they carry the reference's bits
 * (mergenet_b200/segmenter.py does this).  Strictly sequential (one GPU thread replays heapq and the dict orders):
 * the small-image mode the Python reference itself is (practical to ~128 x 256), not the hot path.
 * Returns MN_STATUS_NO_BACKGROUND where the reference raises UnboundLocalError (no class-0 object to prune into).
 * stats4 (optional): heap pops, merges, heap pushes, pruned objects.
 */
int mn_modeb_segment_host(const float* h_log_class, const float* h_log_same, const float* h_log_diff, int num_classes,
                          int num_offsets, int height, int width, const int* offset_list, double object_merge_factor,
                          double merge_logprob_bias, double prune_threshold, long long* h_mask, int* h_object_class,
                          int* n_instances, long long* stats4);

/* ---- tie-exact replay: the reference's C++ segmenter INCLUDING its order among equal priorities ---------------- */
/*
 * c_run_segmentation above reproduces the reference (segment.cc:539-727) wherever the merge order is decided by the
 * priorities, and uses a fixed rule of its own among EXACTLY equal priorities; the reference's order there is an
 * artefact of libstdc++ -- PriorityCompare sees the priority only (segment.h:270-275), so std::push_heap /
 * std::pop_heap decide, fed in std::unordered_map iteration order (segment.cc:650-652).  On inputs whose partition
 * depends on that order (e.g. block-quantized maps) this entry gives the reference's own result: the same edge pass,
 * then ONE GPU thread that replays the loop with GCC 13 libstdc++'s heap and hash-table orders restated literally
 * (mn_exact.cuh, mn_stl_order.h), including the `objects` map whose iteration order numbers the labels
 * (segment.cc:503-515).  h_mask [H][W] and h_object_class [H*W] therefore equal the reference's RAW output arrays
 * (not only up to relabelling).  Sequential by nature: seconds at 256 x 512, for validation and small / medium
 * images, not the hot path.  Arguments as c_run_segmentation (host buffers; h_adj rewritten in place when
 * same_different_bias != 0), with height before width and the batch entries' `clip` flags.
 * stats4 (optional): queue pops, merges, queue pushes, bucket-arena collections.
 * Failures are loud and leave (0, -1) in the outputs: MN_STATUS_Q_POOL (more than 8 E queue entries alive at once),
 * MN_STATUS_PL_POOL (the live bucket arrays exceed 5 E + 32 N words), MN_STATUS_BAD_ARG (shape limits; an offset list
 * that names one pixel pair twice, which the reference leaves undefined), MN_STATUS_INTERNAL (an adjacency list the
 * reference itself would exit(1) on, segment.cc:664-673).
 */
int mn_exact_segment_host(const float* h_class, int num_classes, float* h_adj, int num_offsets, int height, int width,
                          const int* offset_list, int clip, float same_different_bias, float object_merge_factor,
                          float merge_logprob_bias, int* h_mask, int* h_object_class, int* n_instances,
                          long long* stats4);

/* ---- the step after the path: masks back at the image size, COCO run-length encoding ---- */
/*
 * cv2.resize(mask, (out_width, out_height), interpolation=cv2.INTER_NEAREST) for `batch` int32 masks on
 * the device (egs/cityscape/local/segment.py:147-149): out[y][x] = in[min(floor(y * (H / out_h)), H - 1)]
 * [min(floor(x * (W / out_w)), W - 1)], the scale factors and products evaluated in double like OpenCV's
 * resizeNN.  d_in [B][H][W], d_out [B][out_height][out_width].
 */
int mn_resize_masks_nearest_device(const int* d_in, int batch, int height, int width, int* d_out,
                                   int out_height, int out_width, void* stream);
int mn_resize_masks_nearest_host(const int* h_in, int batch, int height, int width, int* h_out,
                                 int out_height, int out_width);
/*
 * cv2.resize(maps, (out_width, out_height)) -- INTER_LINEAR, float32 -- for `planes` map planes
 * (egs/cityscape/local/segment.py:116-123 resizes the class and offset maps to the segmentation size before the
 * segmenter runs).  Planar layout [planes][H][W] -> [planes][out_height][out_width] (the reference moves the channel
 * axis last for cv2 and back: the arithmetic per channel is the same).  Bit-identical to OpenCV 4.13's generic
 * path, which images of 2 or >= 5 channels take (the class / offset maps have 9..81 / 10..16); OpenCV's 1-, 3- and
 * 4-channel paths round differently and are not claimed.
 */
int mn_resize_maps_bilinear_device(const float* d_in, long long planes, int height, int width, float* d_out,
                                   int out_height, int out_width, void* stream);
int mn_resize_maps_bilinear_host(const float* h_in, long long planes, int height, int width, float* h_out,
                                 int out_height, int out_width);
/*
 * COCO run-length encoding of every instance of ONE int32 label mask [H][W] with labels 0..n_instances
 * (egs/cityscape/local/segment.py:165-186: for i in 1..n: maskUtils.encode(asfortranarray(mask == i))):
 * column-major runs (pycocotools rleEncode), written as the compressed ASCII `counts` string
 * (pycocotools rleToString).  Instance i's string is counts[offsets[i-1] .. offsets[i]) (not
 * NUL-terminated); offsets has n_instances + 1 entries.  Returns MN_STATUS_BAD_ARG when counts_capacity
 * is too small (offsets[n] then holds the size needed).  Host buffers in and out; the work runs on
 * the GPU.
 */
int mn_mask_to_coco_rle_host(const int* h_mask, int height, int width, int n_instances,
                             unsigned char* counts, long long counts_capacity, long long* offsets);

/* device time (CUDA events, ms) of the last mn_resize_masks_nearest_host / mn_resize_maps_bilinear_host / mn_mask_to_coco_rle_host call */
float mn_post_last_ms(void);

/* ---- test hooks (parity tests call these through the same library) --------------------------- */
/* Edge pass + record init of ONE image (host buffers in, host buffers out), i.e. what the reference
 * constructor computes (segment.cc:153-232): clp[N*C], cls[N], and per record slot pixel*K+k
 * same/diff/oml/mp plus lo/hi (-1 when the offset leaves the image). */
int mn_debug_edge_dump(int height, int width, int num_classes, int num_offsets,
                       const int* offset_list, const float* h_class, float* h_adj,
                       float same_different_bias, float object_merge_factor,
                       float merge_logprob_bias, float* clp, int* cls, float* same, float* diff,
                       float* oml, float* mp, int* lo, int* hi);
/* Device libm restatements over n consecutive float bit patterns starting at first_bits:
 * which = 0: logf(x); 1: (float)log(1.0 - (double)x); 2: the same_different_bias transform. */
int mn_debug_libm(int which, unsigned first_bits, unsigned n, float bias, float* h_out);
/* times the edge pass alone on `batch` synthetic images (development hook; average ms per launch) */
int mn_debug_edge_bench(int height, int width, int num_classes, int num_offsets, const int* offset_list,
                        int batch, int iters, int clip, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
