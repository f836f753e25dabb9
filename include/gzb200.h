/* gzb200.h -- C ABI of libgzb200.so, the B200-native (sm_100a CUDA) implementation of Guetzli's
 * butteraugli-guided quantisation search.
 *
 * This is the drop-in boundary: every entry point replaces one piece of the reference's
 * accelerator interface (yyamamoto79/guetzli-cuda-opencl; paths relative to the reference root).
 * Plain pointers and sizes only. All pointers are HOST pointers unless stated; each call does its
 * own host<->device copies on the context's stream and returns when the result is on the host.
 *
 * Error convention: every function returns 0 on success or a negative gzb_status; the message is
 * available from gzb_last_error(). The reference logs CUDA errors and carries on
 * (clguetzli/ocu.h:14); this library never falls back to a CPU path -- if no CUDA device can be
 * used, gzb_create fails with GZB_ERR_CUDA.
 *
 * Threading: a gzb_ctx is single-owner (one host thread at a time); different contexts (on the
 * same or different GPUs) may be used concurrently from different threads. No global mutable
 * state except read-only device tables initialised once per device.
 *
 * Numerics: MODE_CPU semantics of the reference (double intermediates, float storage, no early
 * break in the zeroing loop) -- NOT the float-only semantics of the reference's own kernels.
 */
#ifndef GZB200_H_
#define GZB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gzb_ctx gzb_ctx;

typedef enum {
  GZB_OK = 0,
  GZB_ERR_BAD_ARG = -1,
  GZB_ERR_TOO_SMALL = -2, /* width or height < 32: butteraugli is skipped (guetzli/processor.cc:1171) */
  GZB_ERR_CUDA = -3,
  GZB_ERR_STATE = -4,     /* call sequence violated (e.g. zeroing order before StartBlockComparisons) */
  GZB_ERR_UNSUPPORTED = -5
} gzb_status;

/* guetzli::CoeffData (guetzli/processor.h:29-32) */
typedef struct { int idx; float block_err; } gzb_coeff_data;

/* ---- library ---------------------------------------------------------------------------- */
const char* gzb_version(void);
/* Message of the last failure on this context (or of the last failed gzb_create when ctx==NULL). */
const char* gzb_last_error(const gzb_ctx* ctx);
int gzb_device_count(void);

/* ---- comparator life cycle ---------------------------------------------------------------
 * guetzli::ButteraugliComparator(width, height, &rgb, target_distance, stats)
 * (guetzli/butteraugli_comparator.cc:48-58) and ButteraugliComparatorEx, which caches the
 * original's opsin-dynamics image once (clguetzli/clguetzli.cl.cpp:43-69). `rgb` is interleaved
 * sRGB8, width*height*3 bytes, copied. */
int gzb_create(int device, int width, int height, const uint8_t* rgb, float target_distance,
               gzb_ctx** out);
void gzb_destroy(gzb_ctx* ctx);

/* ---- candidate image (device-resident mirror of guetzli::OutputImage) ------------------------
 * Coefficient planes are int16, block-major, ceil(w/8)*ceil(h/8)*64 values per component
 * (OutputImageComponent::coeffs(), guetzli/output_image.h). After gzb_downsample_420 /
 * gzb_set_sampling(ctx, 2) the image is YUV 4:2:0 and the component layouts are those described at
 * gzb_downsample_420 below. */
/* Uploads the q=1 JPEG coefficients of the (4:4:4) input (jpg.components[c].coeffs); the context
 * returns to 4:4:4 if it was 4:2:0. */
int gzb_set_jpeg_coeffs(gzb_ctx* ctx, const int16_t* c0, const int16_t* c1, const int16_t* c2);
/* The same coefficients computed ON THE DEVICE from the context's original image: EncodeRGBToJpeg with
 * the all-ones quantiser (guetzli/jpeg_data_encoder.cc:28-117, guetzli/fdct.cc:28-240; exact integer
 * arithmetic). gzb_get_jpeg_coeffs returns them to the host. */
int gzb_rgb_to_jpeg_coeffs_device(gzb_ctx* ctx);
/* OutputImage::CopyFromJpegData (guetzli/output_image.cc:212-228, 481-492): coeff * quant.
 * Replaces cuCopyFromJpegComponent (clguetzli/cuguetzli.h:138-148). quant: int[3][64]. */
int gzb_copy_from_jpeg(gzb_ctx* ctx, const int* quant192);
/* OutputImage::ApplyGlobalQuantization (guetzli/output_image.cc:349-360, 573-577).
 * Replaces cuApplyGlobalQuantization (clguetzli/cuguetzli.h:150-156). q: int[3][64]. */
int gzb_apply_global_quantization(gzb_ctx* ctx, const int* q192);
/* img.CopyFromJpegData(jpg) with the all-ones matrix followed by img.ApplyGlobalQuantization(q) -- the
 * head of TryQuantMatrix (guetzli/processor.cc:283-284) -- in one pass over the input coefficients. */
int gzb_quantize_from_jpeg(gzb_ctx* ctx, const int* q192);
/* Replaces the candidate's coefficients wholesale (SetCoeffBlock for every block). */
int gzb_set_coeffs(gzb_ctx* ctx, const int16_t* c0, const int16_t* c1, const int16_t* c2);
int gzb_get_coeffs(gzb_ctx* ctx, int16_t* c0, int16_t* c1, int16_t* c2);
/* Sparse update: n records of {block index, component*64+k, new value} (SetCoeffBlock on the
 * touched blocks only; used by the back-end loop, guetzli/processor.cc:867-874). */
int gzb_update_coeffs(gzb_ctx* ctx, const int32_t* block_ix, const uint8_t* idx, const int16_t* val,
                      size_t n);
/* OutputImage::ToSRGB() (guetzli/output_image.cc:642-701): interleaved sRGB8, w*h*3 bytes.
 * Replaces cuComponentsToPixels (clguetzli/cuguetzli.h:157-163). */
int gzb_to_srgb(gzb_ctx* ctx, uint8_t* rgb_out);

/* ---- Comparator interface (guetzli/comparator.h:29-96) ------------------------------------ */
/* Compare(img): butteraugli distance of the resident candidate to the original; the distance map
 * stays on the device. *distance = distmap_aggregate(). Replaces ButteraugliComparatorEx::Compare
 * (clguetzli/clguetzli.cl.cpp:71-152). */
int gzb_compare(gzb_ctx* ctx, float* distance);
/* The same in two halves: begin enqueues the kernels and returns, end waits and reads the distance.
 * Between them the candidate may be entropy-coded (gzb_candidate_* run on a second stream), but not
 * modified. */
int gzb_compare_begin(gzb_ctx* ctx);
int gzb_compare_end(gzb_ctx* ctx, float* distance);
/* distmap(): width*height floats. */
int gzb_get_distmap(gzb_ctx* ctx, float* distmap_out);
/* DistanceOK(target_mul) / ScoreOutputSize(size) / BlockErrorLimit() for the last Compare. */
int gzb_distance_ok(const gzb_ctx* ctx, double target_mul);
double gzb_score_output_size(const gzb_ctx* ctx, int size);
float gzb_block_error_limit(const gzb_ctx* ctx);
/* Image geometry of the context. */
int gzb_image_size(const gzb_ctx* ctx, int* width, int* height);
/* StartBlockComparisons / FinishBlockComparisons (guetzli/butteraugli_comparator.cc:72-83):
 * builds the per-block mask scale (mask_xyz_ at each block's top-left pixel) on the device. */
int gzb_start_block_comparisons(gzb_ctx* ctx);
int gzb_finish_block_comparisons(gzb_ctx* ctx);
/* The side channel of the reference's GPU modes: ButteraugliComparatorEx::
 * imgMaskXyzScaleBlockList ([3] per block) and imgOpsinDynamicsBlockList ([R64 G64 B64] per block)
 * (clguetzli/clguetzli.h:201-217). Either pointer may be NULL. */
int gzb_get_block_lists(gzb_ctx* ctx, float* mask_scale_out, float* opsin_blocks_out);
/* CompareBlock for every 8x8 block of the resident candidate at once (factor 1, comp_mask 7):
 * err_out[block]. (guetzli/butteraugli_comparator.cc:113-163) 4:4:4 candidates only
 * (GZB_ERR_UNSUPPORTED otherwise), like gzb_get_block_lists and gzb_compare_block. */
int gzb_compare_blocks(gzb_ctx* ctx, float* err_out);
/* Comparator::CompareBlock after SwitchBlock(block_x, block_y, 1, 1): the error of one 8x8 block
 * whose candidate coefficients are candidate192 = [Y64 Cb64 Cr64] (dequantised values). One tiny
 * launch per call -- correct but latency-bound; the batched call below is the fast path.
 * (guetzli/butteraugli_comparator.cc:85-163) */
int gzb_compare_block(gzb_ctx* ctx, int block_x, int block_y, const int16_t* candidate192, double* err);
/* Comparator::CompareBlock for any chroma sampling: rgb192 is the candidate's 8x8 window at block
 * (block_x, block_y) of the 8x8 grid as interleaved sRGB8 -- OutputImage::ToSRGB(8*block_x, 8*block_y,
 * 8, 8) (guetzli/output_image.cc:642-701), which the caller renders from its own OutputImage. With
 * SwitchBlock(bx, by, fx, fy) and CompareBlock(img, off_x, off_y, ...) the block is (bx*fx + off_x,
 * by*fy + off_y) (guetzli/butteraugli_comparator.cc:113-122). One tiny launch per call. */
int gzb_compare_block_srgb(gzb_ctx* ctx, int block_x, int block_y, const uint8_t* rgb192, double* err);
/* cuComputeBlockZeroingOrder (clguetzli/cuguetzli.h:30-40) == the per-block loop of
 * Processor::SelectFrequencyMasking over ComputeBlockZeroingOrder (guetzli/processor.cc:376-487,
 * 638-672), MODE_CPU semantics. out: nblocks*192 records, zero-filled, packed from slot 0 in
 * zeroing order, entries with err <= BlockErrorLimit only. Needs gzb_set_jpeg_coeffs (the original
 * coefficients) and gzb_start_block_comparisons. For a 4:2:0 image comp_mask must be 1 (the luma
 * blocks; Cb / Cr of each window are the image's upsampled samples) or 6 (nblocks = the 16x16
 * macro-blocks, ceil(w/16)*ceil(h/16); the error of a trial is the maximum over the macro-block's
 * 8x8 windows, guetzli/processor.cc:426-441). */
int gzb_compute_block_zeroing_order(gzb_ctx* ctx, int comp_mask, gzb_coeff_data* out);
/* Same search, with the candidate packing of SelectFrequencyMasking (guetzli/processor.cc:694-712)
 * done on the device: offsets[nblocks+1] (candidate_coeff_offsets) and, for the records with
 * 0 < err <= BlockErrorLimit in order, cand_idx (candidate_coeffs) / cand_err
 * (candidate_coeff_errors). *n_out = total candidates; the two arrays are filled when cap >= n. */
int gzb_compute_block_zeroing_candidates(gzb_ctx* ctx, int comp_mask, int* offsets, uint8_t* cand_idx,
                                         float* cand_err, size_t cap, size_t* n_out);
/* The same for the blocks [block_begin, block_end) only (a group of GPUs splits the blocks):
 * offsets[block_end - block_begin + 1] are relative to the range. */
int gzb_compute_block_zeroing_candidates_range(gzb_ctx* ctx, int comp_mask, int block_begin, int block_end,
                                               int* offsets, uint8_t* cand_idx, float* cand_err, size_t cap,
                                               size_t* n_out);
/* ---- the candidate as a JPEG, coded on the device ----------------------------------------------
 * OutputImage::SaveToJpegData + WriteJpeg (guetzli/output_image.cc:579-640, guetzli/jpeg_data_writer.cc:
 * 361-553) for the resident candidate: the reference serialises every candidate only to learn its
 * size (guetzli/processor.cc:290-296, 897-903). */
/* BuildDCHistograms / BuildACHistograms (jpeg_data_writer.cc:189-247) of the quantised candidate:
 * raw symbol counts, dc_hist[3][16], ac_hist[3][256]. q192 (may be NULL = the matrix last given to
 * gzb_copy_from_jpeg / gzb_apply_global_quantization) is what the candidate's values are multiples of. */
int gzb_candidate_symbol_histograms(gzb_ctx* ctx, const int* q192, uint32_t* dc_hist48, uint32_t* ac_hist768);
/* The same for a file of ncomp components (1: luma only, in image block order -- what a 4:2:0 candidate
 * whose chroma planes are all zero is written as; 3: the candidate's own MCU order). */
int gzb_candidate_symbol_histograms_n(gzb_ctx* ctx, const int* q192, int ncomp, uint32_t* dc_hist48,
                                      uint32_t* ac_hist768);
/* EncodeScan (jpeg_data_writer.cc:249-359) with the given per-component code tables
 * (dc_code/dc_len[3][16], ac_code/ac_len[3][256]); ncomp is 1 when both chroma planes are all zero
 * (output_image.cc:588), else 3. The scan stays on the device; *scan_bytes is its length before
 * byte stuffing (the last byte padded with ones), *ff_bytes the number of 0xff bytes in it:
 * file size = header + scan_bytes + ff_bytes + 2. */
int gzb_candidate_entropy_code(gzb_ctx* ctx, int ncomp, const uint16_t* dc_code, const uint8_t* dc_len,
                               const uint16_t* ac_code, const uint8_t* ac_len, uint64_t* scan_bytes,
                               uint64_t* ff_bytes);
/* The same when the caller knows the scan's length in bits (it follows from the symbol histograms and the
 * code lengths): saves the round trip between the sizing and the emission pass. Fails with GZB_ERR_STATE
 * if the length does not match the candidate. */
int gzb_candidate_entropy_code_sized(gzb_ctx* ctx, int ncomp, const uint16_t* dc_code, const uint8_t* dc_len,
                                     const uint16_t* ac_code, const uint8_t* ac_len, uint64_t total_bits,
                                     uint64_t* scan_bytes, uint64_t* ff_bytes);
/* Copies the (unstuffed) scan of the last gzb_candidate_entropy_code to the host. */
int gzb_candidate_fetch_scan(gzb_ctx* ctx, uint8_t* out, uint64_t nbytes);
/* The whole file: histograms, ClusterHistograms + code construction (host), scan on the device.
 * input_tables != 0 writes the q tables the way the RGB front end leaves them. *size_out is the file
 * size; the bytes are produced (header + stuffed scan + EOI) only when out != NULL and cap suffices. */
int gzb_write_candidate_jpeg(gzb_ctx* ctx, const int* q192, int input_tables, uint8_t* out, size_t cap,
                             size_t* size_out);
/* ComputeBlockDCTDouble / ComputeBlockIDCTDouble (guetzli/dct_double.cc:47-85), batched: nblocks
 * blocks of 64 doubles, in place. Only the reference's 4:2:0 path uses these transforms. */
int gzb_dct_double(int device, double* blocks, size_t nblocks, int inverse);
/* ComputeBlockErrorAdjustmentWeights (guetzli/butteraugli_comparator.cc:169-233), factor 1.
 * distmap == NULL uses the device-resident map of the last Compare. block_weight: nblocks floats,
 * overwritten (the reference passes a zero vector, guetzli/processor.cc:776-783). */
int gzb_compute_block_error_adjustment_weights(gzb_ctx* ctx, int direction, int max_block_dist,
                                               double target_mul, const float* distmap,
                                               float* block_weight);

/* The same with the sampling factor of the searched component (1: 8x8 blocks, 2: the 16x16 macro-blocks
 * of a 4:2:0 image); block_weight has ceil(w/(8f))*ceil(h/(8f)) entries. */
int gzb_compute_block_error_adjustment_weights_f(gzb_ctx* ctx, int direction, int max_block_dist,
                                                 double target_mul, int factor, const float* distmap,
                                                 float* block_weight);

/* ---- SelectFrequencyBackEnd on the device (guetzli/processor.cc:723-919) ---------------------
 * The adjustment loop's state -- candidate lists, last_indexes, max_block_error, block_weight and the
 * sorted `global_order` of up to 10^7 (block, value) pairs -- lives in HBM. The host keeps only what is
 * sequential by definition: the stopping test on the estimated size over the LAST few hundred steps of
 * the walk. The order is sorted lazily with exactly the element movements of libstdc++'s std::sort (the
 * reference sorts with it and cross-block ties are the norm), see csrc/gzb_backend.cuh. */
typedef struct { int block; float value; } gzb_order_entry;   /* std::pair<int, float> of global_order */
/* What the host walk needs to know about one block (gzb_be_gather). */
typedef struct {
  int last_index;            /* last_indexes[block] */
  unsigned prefix_count;     /* flips of this block in the prefix of the current iteration */
  unsigned long long zmask[3]; /* per component: bit z set = the coefficient at zig-zag position z is non-zero */
  int16_t idx[3][64];        /* quantised indices (coefficient / q), natural order; components outside comp_mask: 0 */
  int16_t requant[3][64];    /* Quantize(input coefficient, q) (quantize.h:24-29): what a "down" step writes; only filled
                                for direction -1 */
} gzb_be_block_state;
/* Starts a pass over the units of comp_mask (8x8 blocks; 16x16 macro-blocks for the chroma pass of a
 * 4:2:0 image): last_indexes and max_block_error are zeroed (processor.cc:757-758). The candidate lists
 * (processor.cc:694-716) are taken from the host arrays, or -- offsets == NULL -- from the lists the last
 * gzb_compute_block_zeroing_candidates call over all blocks of comp_mask left on the device (`total`
 * = its *n_out). */
int gzb_be_begin(gzb_ctx* ctx, int comp_mask, const int* offsets, const uint8_t* cand_idx,
                 const float* cand_err, size_t total);
/* processor.cc:775-828 without the sort: ComputeBlockErrorAdjustmentWeights for rblock = 1, 2, .. until
 * the order is non-empty, then global_order in the reference's arrangement, on the device.
 * *n = global_order.size(), *below = entries with value < below_limit (the partition_point of the
 * first "up" iteration, processor.cc:840-848), *rblock = the radius used (4 if the order stayed empty). */
/* gzb_be_begin for rank `rank` of a group of `world` contexts that each hold the resident candidate lists of
 * their own block range [units * r / world, units * (r + 1) / world) (gzb_compute_block_zeroing_candidates_range):
 * the lists are all-gathered on the devices through `allgather_device` (counts[r] = candidates of rank r,
 * global_offsets[units + 1] = the offsets of the concatenation) and, on the ranks with begin != 0, become the
 * back end's lists; cand_idx_out (optional, sum(counts) bytes) receives the coefficient indices for the host. */
int gzb_be_begin_gathered(gzb_ctx* ctx, int comp_mask, int world, int rank, const int* global_offsets,
                          const uint64_t* counts, int (*allgather_device)(void*, const void*, size_t, void*),
                          void* user, int begin, uint8_t* cand_idx_out);
/* Test hook: cudaMemcpy device to device (a thread-group stand-in for the device all-gather). */
int gzb_test_memcpy_d2d(void* dst, const void* src, size_t nbytes);
int gzb_be_build_order(gzb_ctx* ctx, int direction, double target_mul, float below_limit, uint64_t* n,
                       int* blocks_to_change, uint64_t* below, int* rblock);
/* One step of the lazy std::sort. Pending ranges that end at or before p_set are dropped unsorted (their
 * entries are consumed as a set); the leftmost remaining range is partitioned, the library's way, until
 * it holds at most small_max (<= GZB_BE_MAX_ENTRIES) entries. status 1: entries_out receives that range
 * [*first, *last) -- still to be sorted by the caller with depth budget *depth (std::__introsort_loop's
 * third argument) -- and the range stays pending; 2: the range [*first, *last) has used up its depth
 * budget (std::sort heap-sorts it: fetch, std::partial_sort, store); 3: nothing is pending. */
int gzb_be_select(gzb_ctx* ctx, uint64_t p_set, int small_max, int* status, uint64_t* first,
                  uint64_t* last, int* depth, gzb_order_entry* entries_out);
#define GZB_BE_MAX_ENTRIES 8192   /* entries one select hands over; blocks one gather returns */
#define GZB_BE_MAX_RANGES 16
/* The same for a caller that expects to need the order up to position want_end: up to GZB_BE_MAX_RANGES consecutive
 * short ranges (leftmost first, at most GZB_BE_MAX_ENTRIES entries together, entries back to back in entries_out) in one
 * round trip -- the device goes on to the next pending range as long as one thread block can partition
 * it. status 1: *nranges ranges; 2: ranges[0] needs the heap sort; 3: nothing is pending. */
typedef struct { uint64_t first, last; int depth; int reserved; } gzb_be_range;
int gzb_be_select_ranges(gzb_ctx* ctx, uint64_t p_set, int small_max, uint64_t want_end, int* status,
                         int* nranges, gzb_be_range* ranges, gzb_order_entry* entries_out);
int gzb_be_fetch_order(gzb_ctx* ctx, uint64_t first, gzb_order_entry* out, size_t n);
int gzb_be_store_order(gzb_ctx* ctx, uint64_t first, const gzb_order_entry* in, size_t n);
/* Consumes order[0, p) as a set (processor.cc:854-876 for each entry; the order of the entries does not
 * matter before the first observable step): every block takes as many of its next candidates as it has
 * entries there. ac_hist768 (optional) receives the AC symbol counts of the candidate afterwards, counted
 * as a file of hist_ncomp components; *changed_blocks the number of blocks flipped. blocks / nreq /
 * states_out (optional): a gzb_be_gather in the same round trip. */
int gzb_be_apply_prefix(gzb_ctx* ctx, uint64_t p, int direction, int hist_ncomp, uint32_t* ac_hist768,
                        int* changed_blocks, const int* blocks, int nreq, gzb_be_block_state* states_out);
/* State of nreq (<= GZB_BE_MAX_ENTRIES) blocks for the sequential part of the walk. */
int gzb_be_gather(gzb_ctx* ctx, const int* blocks, int nreq, int direction, gzb_be_block_state* states_out);
/* End of an iteration: the flips of the sequential walk (unit index, coefficient index 64*c + k, value;
 * each advances last_indexes by direction), max_block_error += block_weight * val_threshold * direction
 * (processor.cc:893-895), and the candidate's samples are re-rendered for the next Compare. */
int gzb_be_finish_iteration(gzb_ctx* ctx, const int32_t* blocks, const uint8_t* cidx, const int16_t* val,
                            size_t n, int direction, float val_threshold);
/* IsGrayscale(jpg) of the input (processor.cc:921-929): *gray = 1 if both chroma planes are all zero. */
int gzb_input_is_gray(gzb_ctx* ctx, int* gray);
/* Counters: selects = gzb_be_select calls, levels = partitions run on the device, host_ranges = short
 * ranges finished by the caller. */
int gzb_be_stats(const gzb_ctx* ctx, unsigned long long* selects, unsigned long long* levels);
/* Test hook: sorts n entries through the same device path (partitions on the device down to small_max,
 * short ranges finished with the host's restatement of introsort) and checks nothing else. The result
 * must equal std::sort's permutation. prefix > 0: additionally treats [0, prefix) as a set first. */
int gzb_test_device_sort(gzb_ctx* ctx, gzb_order_entry* entries, size_t n, size_t prefix, int small_max);
/* Test hook: makes `entries` the context's order (no candidate lists behind it) and restarts the lazy sort. */
int gzb_be_test_load_order(gzb_ctx* ctx, const gzb_order_entry* entries, size_t n);
/* The same with a depth budget other than 2 * log2(n): a small one drives the sort into introsort's heap-sort fallback. */
int gzb_be_test_load_order_depth(gzb_ctx* ctx, const gzb_order_entry* entries, size_t n, int depth);
int gzb_test_device_sort_depth(gzb_ctx* ctx, gzb_order_entry* entries, size_t n, size_t prefix, int small_max, int depth);
/* Host checker for it: the restated introsort over the whole array with the given depth budget. */
void gzb_test_exact_sort_depth(int* first, float* second, size_t n, int depth);

/* ---- YUV 4:2:0 (Params::try_420 / force_420; guetzli/processor.cc:986-1016) ------------------
 * Processor::DownsampleImage + OutputImage::SaveToJpegData on the q=1 input (guetzli/processor.cc:
 * 109-116, 994-997; guetzli/output_image.cc:496-640; guetzli/preprocess_downsample.cc:157-281, default
 * DownsampleConfig): the input coefficients become those of the sharpened / blurred, 2x2-averaged
 * chroma planes (double-precision DCT, guetzli/dct_double.cc) and the context switches to 4:2:0:
 * component 0 keeps its blocks, laid out MCU-padded (2*ceil(w/16) x 2*ceil(h/16), padding blocks =
 * {DC of the predecessor, 0...}), components 1 and 2 have ceil(w/16) x ceil(h/16) blocks. From here on
 * gzb_copy_from_jpeg / gzb_apply_global_quantization / gzb_set_coeffs / gzb_get_coeffs /
 * gzb_update_coeffs (block index = index in the component's own layout) / gzb_to_srgb / gzb_compare /
 * the zeroing search (comp_mask 1: luma blocks, 6: macro-blocks) / gzb_candidate_* work on the
 * factor-2 image (UpdatePixelsForBlock's fancy upsampling, guetzli/output_image.cc:147-210).
 * gzb_set_jpeg_coeffs returns the context to 4:4:4. */
int gzb_downsample_420(gzb_ctx* ctx);
/* Switches the context between 4:4:4 (chroma_factor 1) and 4:2:0 (2) without touching any data: a
 * caller that owns the OutputImage (the Comparator adaptor) then sends coefficients in the layout above
 * with gzb_set_coeffs / gzb_set_jpeg_coeffs_420. Invalidates the resident coefficients on a change. */
int gzb_set_sampling(gzb_ctx* ctx, int chroma_factor);
/* JPEGData of a 4:2:0 image (jpg.components[c].coeffs; luma 2*ceil(w/16) x 2*ceil(h/16) blocks, chroma
 * ceil(w/16) x ceil(h/16)) as the input coefficients; switches the context to 4:2:0. */
int gzb_set_jpeg_coeffs_420(gzb_ctx* ctx, const int16_t* c0, const int16_t* c1, const int16_t* c2);
/* Blocks per row / column of component comp's coefficient arrays and its sampling factor. */
int gzb_component_dims(const gzb_ctx* ctx, int comp, int* blocks_w, int* blocks_h, int* factor);
/* The input coefficients (jpg.components[c].coeffs after SaveToJpegData); NULL skips a component. */
int gzb_get_jpeg_coeffs(gzb_ctx* ctx, int16_t* c0, int16_t* c1, int16_t* c2);

/* ---- stage entry points (the reference's cu* free functions; host planes in, host planes out) */
/* cuOpsinDynamicsImage (clguetzli/cuguetzli.h:19-21): linear rgb planes in place -> XYB. */
int gzb_opsin_dynamics_image(int device, float* r, float* g, float* b, size_t xsize, size_t ysize);
/* cuDiffmapOpsinDynamicsImage (clguetzli/cuguetzli.h:23-28). step must be 3. */
int gzb_diffmap_opsin_dynamics_image(int device, float* result, const float* r, const float* g,
                                     const float* b, const float* r2, const float* g2,
                                     const float* b2, size_t xsize, size_t ysize, size_t step);
/* butteraugli::Blur (butteraugli.cc:100-148) / cuBlurEx (clguetzli/cuguetzli.h:73-75). */
int gzb_blur(int device, float* plane, size_t xsize, size_t ysize, double sigma,
             double border_ratio);
/* butteraugli::Mask (third_party/butteraugli/butteraugli/butteraugli.cc:1505-1566): replaces cuMask
 * (clguetzli/cuguetzli.h:42-47; dispatcher clguetzli/clbutter_comparator.cpp:1651-1666). Six XYB planes in
 * (xsize*ysize floats each), mask and mask_dc planes out. */
int gzb_mask(int device, float* mask_r, float* mask_g, float* mask_b, float* maskdc_r, float* maskdc_g,
             float* maskdc_b, size_t xsize, size_t ysize, const float* r, const float* g, const float* b,
             const float* r2, const float* g2, const float* b2);
/* Standalone butteraugli of two sRGB8 images (interleaved): distance and optional diffmap. */
int gzb_butteraugli_srgb(int device, const uint8_t* rgb0, const uint8_t* rgb1, int width,
                         int height, float* distance, float* diffmap_out);

/* Zeroes the device-resident distance map (the back end's first "up" iteration runs
 * ComputeBlockErrorAdjustmentWeights on an all-zero map, guetzli/processor.cc:777-780). */
int gzb_clear_distmap(gzb_ctx* ctx);

/* ---- whole encoder: guetzli::Process(params, stats, rgb, w, h, &out) ------------------------
 * (guetzli/processor.cc:1157-1185 with the default Params, guetzli/processor.h:34-42): RGB ->
 * q=1 coefficients -> SelectQuantMatrix -> SelectFrequencyMasking -> best JPEG. The search driver
 * and the Huffman writer run on host threads; every butteraugli evaluation, IDCT/quantisation pass
 * and the zeroing-order search run on the B200. Output bytes equal the CPU reference's.
 * *jpeg_out is malloc'ed (release with gzb_free). host_threads <= 0 picks min(16, cores).
 * trace_out, if non-NULL, receives a malloc'ed copy of the verbose trace in the reference's
 * GUETZLI_LOG format (guetzli/processor.cc:324-331, 905-913). */
typedef struct {
  int num_iterations, num_iterations_up, num_iterations_down;  /* ProcessStats counters */
  int num_compares, num_jpeg_writes, num_entropy_code_builds;
  double total_wall_ms, host_frontend_ms, host_quant_ms, host_write_ms;
  double compare_wall_ms, device_compare_ms, zeroing_wall_ms, device_zeroing_ms, backend_wall_ms;
  double write_hist_ms, write_code_ms, write_encode_ms, write_stitch_ms;   /* parts of host_write_ms */
  double be_weights_ms, be_order_ms, be_walk_ms, be_update_ms, create_ms;  /* parts of the back end */
  double be_codes_ms, be_sort_ms;        /* inside be_walk_ms: entropy-code rebuilds, lazy sort */
  unsigned long long be_steps;           /* coefficients flipped by the back end */
  double prepare_ms, run_ms;             /* gzb_encoder_create / gzb_encoder_run wall time */
  unsigned long long h2d_bytes, d2h_bytes; /* host<->device traffic of the whole encode */
  double final_score;
  float final_distance;
  unsigned long long launches;
  unsigned long long be_prefix_steps;    /* of be_steps: applied block-parallel in the silent prefix */
  double device_write_ms;                /* candidate files coded on the device (histograms, codes, scan, fetch) */
  double search_wall_ms, trial_host_ms, trial_device_ms;  /* SelectQuantMatrix phase; host/device legs of its trials */
  int search_rounds, search_trials;      /* SelectQuantMatrix: exchange rounds / trials evaluated by the group */
  double downsample_ms;                  /* YUV420 passes: DownsampleImage + SaveToJpegData on the device */
  unsigned long long be_selects, be_levels;  /* back end: lazy-sort kernel launches / partitions run on the device */
  unsigned long long be_host_ranges;     /* back end: short ranges of the order finished on the host */
  double be_lazy_ms;                     /* inside be_sort_ms: host finishing of the short ranges */
  unsigned long long num_fine_bdm_compares;  /* of num_compares: BlockDiffMap recomputed only around the flipped blocks */
  double be_select_ms, be_gather_ms, be_pool_ms;  /* inside be_walk_ms: lazy-sort round trips, block-state round trips
                                                     (the first one consumes the prefix), parallel entropy-code rebuilds */
} gzb_encode_stats;
int gzb_encode_rgb(int device, const uint8_t* rgb, int width, int height, float butteraugli_target,
                   int host_threads, uint8_t** jpeg_out, size_t* jpeg_size, gzb_encode_stats* stats,
                   char** trace_out);
/* guetzli::Process with Params::try_420 / Params::force_420 (guetzli/processor.h:34-42; the passes of
 * ProcessJpegData, guetzli/processor.cc:986-1016): try_420 runs the 4:4:4 pass and then the YUV420
 * pass (downsampling, quant search from score 0, frequency masking of luma then of chroma with early
 * stop) unless the image is grey; force_420 runs the YUV420 pass only. */
int gzb_encode_rgb_params(int device, const uint8_t* rgb, int width, int height, float butteraugli_target,
                          int try_420, int force_420, int host_threads, uint8_t** jpeg_out, size_t* jpeg_size,
                          gzb_encode_stats* stats, char** trace_out);
/* A batch of n images on ONE GPU (BASELINE configs[3], one rank's share) with `inflight` encodes running
 * concurrently, each on host_threads_per_encode host threads (<= 0: the cores divided by inflight) and its
 * own device context: the sequential host phases of one encode overlap the kernels of the others
 * (measured: 1.7x the one-at-a-time throughput with three in flight). rgb / width / height / jpeg_out /
 * jpeg_size / stats / status are arrays of n; jpeg_out[i] is malloc'ed (gzb_free). Each result equals
 * that of gzb_encode_rgb_params on its own. Returns the first failing image's code, or GZB_OK. */
int gzb_encode_rgb_batch(int device, int n, const uint8_t* const* rgb, const int* width, const int* height,
                         float butteraugli_target, int try_420, int force_420, int inflight,
                         int host_threads_per_encode, uint8_t** jpeg_out, size_t* jpeg_size,
                         gzb_encode_stats* stats, int* status);
/* The same encoder in two steps, so that a caller can separate "inputs resident in HBM" from the
 * search: create uploads the image, computes its opsin-dynamics image and the q=1 coefficients;
 * run performs the search (once per encoder). */
typedef struct gzb_encoder gzb_encoder;
int gzb_encoder_create(int device, const uint8_t* rgb, int width, int height, float butteraugli_target,
                       int host_threads, gzb_encoder** out);
int gzb_encoder_run(gzb_encoder* enc, uint8_t** jpeg_out, size_t* jpeg_size, gzb_encode_stats* stats,
                    char** trace_out);
gzb_ctx* gzb_encoder_context(gzb_encoder* enc);
/* Params::try_420 / force_420 for a two-step encoder; call before gzb_encoder_run. */
int gzb_encoder_set_params(gzb_encoder* enc, int try_420, int force_420);
/* Multi-GPU, one large image: the encoders of a group -- one per GPU, one per
 * process, all created from the SAME image and target -- share the work of one encode.
 *   * SelectQuantMatrix (guetzli/processor.cc:310-372): the independent TryQuantMatrix candidates
 *     the generator may ask for next are evaluated one per rank; after each round the ranks
 *     all-gather {distance, jpg_size} and every rank replays the reference's sequential decisions,
 *     so the visited sequence, the trace and the chosen matrix equal the single-GPU run's.
 *   * The block-zeroing search (processor.cc:638-672) is split by block range and its candidate
 *     lists are all-gathered.
 *   * The back end (processor.cc:723-919) is one sequential walk: rank 0 runs it and returns the
 *     JPEG; the other ranks return from gzb_encoder_run with *jpeg_size == 0.
 *   * With try_420 / force_420 (processor.cc:986-1016) every rank downsamples its own copy; the quant
 *     searches of both passes and the zeroing searches of the 4:4:4 pass and of the luma pass are
 *     shared as above. The chroma pass starts from the candidate rank 0's luma back end leaves
 *     behind, which the other ranks do not have: rank 0 runs its zeroing search alone.
 * A rank whose work fails still enters every exchange and reports the failure there, so all ranks
 * return an error together instead of the others waiting inside the collective.
 * `allgather` must gather `nbytes` from every rank into recv[world*nbytes] in rank order (e.g.
 * ncclAllGather / torch.distributed.all_gather over NCCL) and return 0; it is called the same
 * number of times with the same sizes on every rank. Call before gzb_encoder_run. */
typedef int (*gzb_allgather_fn)(void* user, const void* send, size_t nbytes, void* recv);
int gzb_encoder_set_group(gzb_encoder* enc, int rank, int world, gzb_allgather_fn allgather, void* user);
/* Optional second exchange function for the group, for the one bulky exchange (the zeroing candidates: 5 bytes
 * per candidate, tens of MB at 12 MPix): an all-gather of DEVICE memory -- nbytes at d_send on this rank's GPU
 * into d_recv (world * nbytes, rank order) -- e.g. ncclAllGather on the caller's communicator or
 * torch.distributed.all_gather_into_tensor on tensors that alias the pointers. The library has synchronised
 * its own stream before the call; the function returns when d_recv is complete. With it the candidate lists go
 * from GPU to GPU over NVLink and never visit the host; without it they are staged through `allgather`. `user`
 * is the pointer given to gzb_encoder_set_group. Call after gzb_encoder_set_group. */
typedef int (*gzb_allgather_device_fn)(void* user, const void* d_send, size_t nbytes, void* d_recv);
int gzb_encoder_set_group_device(gzb_encoder* enc, gzb_allgather_device_fn allgather_device);
void gzb_encoder_destroy(gzb_encoder* enc);
const char* gzb_encode_last_error(void);
void gzb_free(void* p);
/* guetzli::ButteraugliScoreForQuality (guetzli/quality.cc:76-85). */
double gzb_butteraugli_score_for_quality(double quality);
/* guetzli::EncodeRGBToJpeg with the all-ones quantiser (guetzli/jpeg_data_encoder.cc:66-136):
 * block-major int16 coefficient planes of ceil(w/8)*ceil(h/8)*64 values each. Host code. */
int gzb_rgb_to_jpeg_coeffs(const uint8_t* rgb, int width, int height, int16_t* c0, int16_t* c1,
                           int16_t* c2);
/* Serialises coefficient planes (dequantised values, multiples of q) to a JPEG byte stream exactly
 * as OutputImage::SaveToJpegData + WriteJpeg do (guetzli/output_image.cc:579-640,
 * guetzli/jpeg_data_writer.cc:540-553). input_tables != 0 writes the q tables the way the RGB
 * front end leaves them (three tables with index 0). Returns the size; copies if it fits cap. */
long gzb_write_jpeg(const int16_t* c0, const int16_t* c1, const int16_t* c2, int width, int height,
                    const int* q192, int input_tables, int host_threads, uint8_t* out, long cap);

/* ---- parity/debug: device intermediates of the last gzb_compare --------------------------- */
/* name in {"xyb0","xyb1","mhic0","mhic1","edge_map","block_dc","block_ac","combined_sqrt",
 * "diffmap","mask_front"}; copies min(cap, size) floats; *n_out = size in floats. */
int gzb_debug_fetch(gzb_ctx* ctx, const char* name, float* out, size_t cap, size_t* n_out);

/* ---- per-kernel device timing (CUDA events on the context's stream around every launch) ---- */
/* Off by default. While on, every kernel launch of this context is bracketed by an event pair and
 * accumulated per kernel name. */
int gzb_profile_enable(gzb_ctx* ctx, int on);
/* Number of kernel classes; name / accumulated ms / launch count of class i. */
int gzb_profile_count(void);
const char* gzb_profile_name(int i);
int gzb_profile_get(gzb_ctx* ctx, int i, double* ms, unsigned long long* launches);
int gzb_profile_reset(gzb_ctx* ctx);
/* Host->device and device->host bytes moved by this context since creation. */
/* Measurement aid: the device's double-precision rate WITHOUT fused multiply-add (DADD + DMUL streams on all
 * SMs), in Gflop/s. The search kernels are built with -fmad=false to match the reference's arithmetic, so this
 * -- not the data-sheet FMA rate -- is their arithmetic ceiling. */
int gzb_measure_fp64_peak(int device, double* gflops);
/* Test hook (host arithmetic only): the FMA-based quotient used by the opsin gamma against IEEE division for every
 * float argument in [0, 1024]; returns the number of mismatches (must be 0). */
unsigned long long gzb_test_gamma_division(void);
int gzb_get_transfer_bytes(const gzb_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h);

/* ---- timing of the last call on this context (CUDA events on the context's stream) -------- */
/* Device milliseconds spent by the kernels of the last gzb_compare / zeroing call. */
float gzb_last_device_ms(const gzb_ctx* ctx);
/* Blocks of the last zeroing search with more than 16 candidates and two equal ordering keys: their input order
 * (guetzli/processor.cc:410-412) is std::sort's own arrangement, reproduced by the restated introsort. */
unsigned gzb_last_zeroing_tie_blocks(const gzb_ctx* ctx);
/* Number of kernels launched by this context since creation (bench `gpu_launches`). */
unsigned long long gzb_launch_count(const gzb_ctx* ctx);
/* Number of Compares of this context that were incremental: between two Compares separated only by
 * gzb_update_coeffs calls touching few blocks (the back end's iterations), every stage of the pipeline
 * recomputes only the 32x32-pixel tiles the changed blocks can reach (bounded supports of the blurs and
 * block transforms); all other values are still in the context's buffers. The results are bit-identical
 * to a full Compare. GZB_NO_INCREMENTAL=1 in the environment disables it. */
unsigned long long gzb_incremental_compare_count(const gzb_ctx* ctx);
/* Compares whose BlockDiffMap cells were recomputed only around the blocks flipped since the previous Compare
 * (every other stage ran in full): the back end's "down" iterations flip 1-5 % of the blocks. */
unsigned long long gzb_fine_bdm_compare_count(const gzb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GZB200_H_ */
