/*
 * rdp_b200_backend.h -- the thin C-ABI between the C++ host (command processor, renderer, video interface)
 * and the CUDA runtime.  It replaces the Granite/Vulkan layer the reference renderer is written against:
 * every entry point below corresponds to a group of Granite call sites in the reference.
 *
 *   buffers / RDRAM mirror   <- Vulkan::Device::create_buffer / map_host_buffer / copy_buffer
 *                               (parallel-rdp/rdp_device.cpp:54-131,245-261; rdp_renderer.cpp:603-658)
 *   pass staging + submit    <- RenderBuffersUpdater::upload            (rdp_renderer.cpp:1592-1632)
 *   rdpb_launch_front        <- Renderer::update_tmem_instances         (rdp_renderer.cpp:1634-1668)
 *                               + CPU-side decode in the reference      (rdp_device.cpp:268-419, rdp_renderer.cpp:910-925)
 *   rdpb_launch_span_bin     <- submit_span_setup_jobs + submit_tile_binning_combined (rdp_renderer.cpp:1670-1700,1832-1935)
 *   rdpb_launch_tiles        <- submit_rasterization + submit_depth_blend (rdp_renderer.cpp:1721-1830,2086-2221)
 *   rdpb_launch_vi_*         <- VideoInterface::*_stage                 (video_interface.cpp:405-999)
 *   rdpb_signal / rdpb_wait  <- Vulkan::Fence + timeline                (rdp_renderer.cpp:2404-2466, rdp_device.cpp:1053-1080)
 *
 * Barriers of the reference (rdp_renderer.cpp:2251,2259,2271,2347) become CUDA stream order.
 * Two implementations exist: the product (CUDA, parallel-rdp_b200/csrc/backend_cuda.cu) and a test-only CPU
 * build of the same device functions (tests/hostsim), selected at link time.  The product never falls back.
 */
#ifndef RDP_B200_BACKEND_H_
#define RDP_B200_BACKEND_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rdpb_device rdpb_device;

/* Host-visible staging area for the pass being recorded.  Pointers stay valid until the pass is uploaded. */
typedef struct rdpb_pass_buffers
{
	uint32_t *words;   /* raw RDP command words of the primitives, capacity rdpb::kMaxPassWords */
	void *records;     /* rdpb::PrimRecord[kMaxPrimsPerPass] */
	void *states;      /* rdpb::PassState[kMaxStatesPerPass] */
	void *uploads;     /* rdpb::TmemUpload[kMaxUploadsPerPass] */
} rdpb_pass_buffers;

typedef struct rdpb_pass_counts
{
	uint32_t num_words;
	uint32_t num_prims;
	uint32_t num_states;
	uint32_t num_uploads;
	uint32_t num_spans;
} rdpb_pass_counts;

/* Returns NULL (and logs to stderr) when no usable CUDA device exists. */
rdpb_device *rdpb_device_create(int cuda_ordinal, uint32_t rdram_size, uint32_t hidden_rdram_size);
void rdpb_device_destroy(rdpb_device *dev);
const char *rdpb_backend_name(void);
const char *rdpb_last_error(rdpb_device *dev);

/* RDRAM / hidden RDRAM / TMEM mirrors.  Copies are asynchronous on the device's stream unless `sync`. */
int rdpb_register_host_memory(rdpb_device *dev, void *ptr, size_t size);
int rdpb_unregister_host_memory(rdpb_device *dev, void *ptr);
int rdpb_rdram_write(rdpb_device *dev, uint32_t offset, const void *src, uint32_t size, int sync);
int rdpb_rdram_read(rdpb_device *dev, uint32_t offset, void *dst, uint32_t size, int sync);
/* Asynchronous read-back of `num_ranges` {offset, size} pairs into host_base + offset on a side stream, behind all work
 * submitted so far, with a timeline point after the copies: rdpb_wait(value) => the bytes are in host memory. */
int rdpb_rdram_read_fenced(rdpb_device *dev, void *host_base, const uint32_t *ranges, uint32_t num_ranges, uint64_t value);
int rdpb_hidden_write(rdpb_device *dev, uint32_t offset, const void *src, uint32_t size, int sync);
int rdpb_hidden_read(rdpb_device *dev, uint32_t offset, void *dst, uint32_t size, int sync);
int rdpb_hidden_fill(rdpb_device *dev, uint8_t value);
/* TMEM (4 KiB) lives in host-mapped pinned memory: the pointer is coherent whenever the stream is idle. */
void *rdpb_tmem_host_ptr(rdpb_device *dev);
void *rdpb_rdram_device_ptr(rdpb_device *dev);
void *rdpb_hidden_device_ptr(rdpb_device *dev);

/* Brackets a batch of submission calls (a whole pass, a scan-out + its read-back): the CUDA backend admits only a few
 * such batches at a time per GPU, so that the command threads of many processors do not stampede the driver. */
void rdpb_submit_begin(rdpb_device *dev);
void rdpb_submit_end(rdpb_device *dev);

/* Pass recording and submission. */
int rdpb_pass_begin(rdpb_device *dev, rdpb_pass_buffers *out);
/* Uploads records / states / texture-load descriptors of the pass from the staging area; command words follow as
 * one rdpb_pass_upload_words() call per contiguous run (`src` is staging memory or caller-owned "borrowed" memory). */
int rdpb_pass_upload(rdpb_device *dev, const rdpb_pass_counts *counts);
int rdpb_pass_upload_words(rdpb_device *dev, const uint32_t *src, uint32_t num_words, uint32_t dst_word_offset);
/* Front end of a pass in one launch: replay `num_uploads` texture loads into TMEM snapshots (when update_tmem & 1),
 * decode `num_prims` primitives from their raw command words and allocate their span jobs.  update_tmem & 2: the
 * pass's texture loads may read memory the previous pass renders to -- the CUDA backend runs set-up work ahead of the
 * previous pass's tile kernel unless told so.  update_tmem & 4: the pass samples TMEM (the CUDA backend keeps two sets
 * of TMEM snapshots and refreshes the one this pass uses when it is out of date even if bit 0 is clear).  update_tmem & 8: a
 * native rdpb_launch_span_bin follows for this pass (it then closes the pass's set-up side; otherwise this call does). */
int rdpb_launch_front(rdpb_device *dev, int update_tmem, uint32_t num_uploads, uint32_t num_prims);
int rdpb_launch_span_bin(rdpb_device *dev, const void *fb_info /* rdpb::FbInfo */);
/* `summary` (rdpb::PassSummary) describes what the pass's states have in common; the backend uses it to pick a
 * specialised instantiation of the tile kernel (NULL: the generic one). */
int rdpb_launch_tiles(rdpb_device *dev, const void *fb_info /* rdpb::FbInfo */, const void *summary /* rdpb::PassSummary */);
/* Name of the tile-kernel instantiation chosen for the most recent pass (diagnostics / tests). */
const char *rdpb_last_tile_variant(rdpb_device *dev);
int rdpb_pass_end(rdpb_device *dev);

/* VI scan-out.  `frame` is rdpb::ViFrame (built by the host from the VI registers), `lines` rdpb::ViLine[288].
 * Two kernels: filter (AA / dither reconstruction / divot straight from RDRAM) and scale (+gamma, field merge).
 * The result is RGBA8 in device memory, double-buffered so the previous field can be merged. */
int rdpb_launch_vi_scanout(rdpb_device *dev, const void *frame /* rdpb::ViFrame */, const void *lines /* rdpb::ViLine[288] */);
void *rdpb_vi_output_device_ptr(rdpb_device *dev);
int rdpb_vi_read_output(rdpb_device *dev, void *dst, uint32_t width, uint32_t height, int sync);
/* Asynchronous read-back on a side stream + a timeline point after it: rdpb_wait(fence_value) => pixels are in dst. */
int rdpb_vi_read_output_fenced(rdpb_device *dev, void *dst, uint32_t width, uint32_t height, uint64_t fence_value);

/* Synchronisation. */
int rdpb_signal(rdpb_device *dev, uint64_t value);   /* records a point in stream order */
int rdpb_wait(rdpb_device *dev, uint64_t value);     /* blocks the calling thread until that point completed */
int rdpb_poll(rdpb_device *dev, uint64_t value);     /* non-blocking: 1 = that point has completed, 0 = still pending */
int rdpb_sync(rdpb_device *dev);
void *rdpb_stream(rdpb_device *dev);                 /* cudaStream_t for interop/benchmark timing */

/* ---- Internal upscaling (replaces rdp_renderer.cpp:545-600 buffers, :1943-2084 domain updates) ------------------
 * The upscaled domain is S*S "slices" of RDRAM and hidden RDRAM (S = 1 << scale_log2, 1..3) plus a native-sized
 * shadow of RDRAM.  A pass that is rendered upscaled is submitted twice: at native resolution as always, bracketed by
 * rdpb_launch_upscale_sync(PRE / POST), then again through rdpb_launch_span_bin / rdpb_launch_tiles with an FbInfo whose
 * scale_log2 is set and whose width / height / tile counts are in upscaled pixels. */
int rdpb_upscaling_init(rdpb_device *dev, uint32_t scale_log2);
/* Super-sampled read-back (replaces update_upscaled_domain_resolve.comp + clear_super_sampled_write_mask.comp,
 * rdp_renderer.cpp:2037-2060, 2296-2334): the pass is submitted upscaled only, with FbInfo::write_mask set, as
 * PRE, span_bin + tiles (upscaled), RESOLVE; RESOLVE averages the samples of every native pixel the pass stored to
 * into native RDRAM and clears the write mask for the next pass.  RESOLVE_DITHER adds the ordered dither for 16-bit
 * colour images. */
enum { RDPB_UPSCALE_SYNC_PRE = 0, RDPB_UPSCALE_SYNC_POST = 1, RDPB_UPSCALE_SYNC_RESOLVE = 2, RDPB_UPSCALE_SYNC_RESOLVE_DITHER = 3 };
int rdpb_launch_upscale_sync(rdpb_device *dev, const void *native_fb_info, int stage);
/* Blocking read-back of the whole upscaled domain: S*S images of rdram_size bytes, S*S images of rdram_size / 2 bytes. */
int rdpb_upscaled_read(rdpb_device *dev, void *rdram_dst, void *hidden_dst);

/* The scan-out runs on its own stream; this makes the compute stream (rdpb_stream) wait for the one in flight, for
 * callers that consume rdpb_vi_output_device_ptr() in compute-stream order. */
int rdpb_vi_join(rdpb_device *dev);

/* Instrumentation: number of kernels launched since creation. */
uint64_t rdpb_launch_count(rdpb_device *dev);

/* ---- Measurement plumbing (bench.py / profiling only; not used by the host renderer) ------------------------ */
enum
{
	RDPB_KERNEL_FRONT = 0, /* texture-load replay + primitive set-up, one launch */
	RDPB_KERNEL_FILL_PASS = 1, /* per-pixel kernel that replaces the tile kernel on small fill-only passes */
	RDPB_KERNEL_SPAN_BIN = 2,
	RDPB_KERNEL_TILES = 3,
	RDPB_KERNEL_VI_FILTER = 4,
	RDPB_KERNEL_VI_SCALE = 5,
	RDPB_KERNEL_TILES_UPSCALED = 6,    /* the tile kernel of an upscaled pass */
	RDPB_KERNEL_SPAN_BIN_UPSCALED = 7,
	RDPB_KERNEL_UPSCALE_SYNC = 8,      /* PRE / POST / RESOLVE of the upscaled domain */
	RDPB_KERNEL_VI_DOWNSCALE = 9,      /* halving steps (+ bob) of an upscaled scan-out */
	RDPB_KERNEL_COUNT = 10
};

/* Per-kernel device time from CUDA events recorded around every launch on the launching stream. */
int rdpb_profile_enable(rdpb_device *dev, int on, uint32_t max_launches);
int rdpb_profile_read(rdpb_device *dev, int kernel, uint32_t *launches, double *total_ms);
/* out8[0] = non-empty tiles processed by the tile kernel, out8[1] = shaded (pixel, primitive) pairs. */
int rdpb_counters_read(rdpb_device *dev, uint64_t *out8, int reset);
/* Record the GPU work of everything submitted between begin/end while keeping its inputs resident in device
 * memory; rdpb_record_mark() separates steps (frames).  rdpb_replay_step() re-issues one step's kernel launches
 * from those resident inputs: no host->device traffic at all. */
int rdpb_record_begin(rdpb_device *dev, size_t arena_bytes);
int rdpb_record_mark(rdpb_device *dev);
int rdpb_record_end(rdpb_device *dev);
uint32_t rdpb_record_steps(rdpb_device *dev);
int rdpb_replay_step(rdpb_device *dev, uint32_t step);
/* Several independent streams on one GPU (BASELINE: "1/2/4/8 streams"): re-issue steps first .. first+count-1 of every
 * device, round-robin per step, so the streams' kernels overlap on the SMs.  Step numbers wrap per device. */
int rdpb_replay_interleaved(rdpb_device *const *devs, uint32_t num_devs, uint32_t first_step, uint32_t count);

#ifdef __cplusplus
}
#endif
#endif
