/*
 * rdp_b200.h -- public C-ABI of librdp_b200.so: the drop-in boundary for emulator / FFI integrations.
 *
 * One entry point per public method of the reference's RDP::CommandProcessor
 * (parallel-rdp/rdp_device.hpp:111-197; behaviour: parallel-rdp/rdp_device.cpp).  Plain pointers and sizes
 * only; no C++ or torch types cross this boundary.  See INTEGRATION.md for the binding a maintainer adds.
 *
 * Ownership: `rdram_ptr` is borrowed for the processor's lifetime (reference: integration_example.cpp:459-505);
 * hidden RDRAM and TMEM are owned by the processor; pointers returned by begin_read_* / get_tmem stay owned by
 * the processor and are valid while it is idle.
 * Errors: no error codes (as in the reference); creation failure => rdp_b200_device_is_supported() == 0 and a
 * message on stderr.  The library never falls back to a CPU renderer.
 */
#ifndef RDP_B200_H_
#define RDP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rdp_b200_processor rdp_b200_processor;

/* CommandProcessorFlagBits, rdp_device.hpp:43-53 */
enum
{
	RDP_B200_FLAG_HOST_VISIBLE_HIDDEN_RDRAM = 1 << 0,
	RDP_B200_FLAG_HOST_VISIBLE_TMEM = 1 << 1,
	RDP_B200_FLAG_UPSCALING_2X = 1 << 2,
	RDP_B200_FLAG_UPSCALING_4X = 1 << 3,
	RDP_B200_FLAG_UPSCALING_8X = 1 << 4,
	RDP_B200_FLAG_SUPER_SAMPLED_READ_BACK = 1 << 5,
	RDP_B200_FLAG_SUPER_SAMPLED_DITHER = 1 << 6,
	/* Extension: see "RDRAM coherency" below (rdp_b200_mark_rdram_written). */
	RDP_B200_FLAG_EXPLICIT_RDRAM_SYNC = 1 << 16
};

/* VIRegister, rdp_common.hpp:324-341 */
enum
{
	RDP_B200_VI_CONTROL = 0, RDP_B200_VI_ORIGIN, RDP_B200_VI_WIDTH, RDP_B200_VI_INTR, RDP_B200_VI_V_CURRENT_LINE,
	RDP_B200_VI_TIMING, RDP_B200_VI_V_SYNC, RDP_B200_VI_H_SYNC, RDP_B200_VI_LEAP, RDP_B200_VI_H_START,
	RDP_B200_VI_V_START, RDP_B200_VI_V_BURST, RDP_B200_VI_X_SCALE, RDP_B200_VI_Y_SCALE
};

/* ScanoutOptions, video_interface.hpp:31-81 (Vulkan export fields dropped). */
typedef struct rdp_b200_scanout_options
{
	uint32_t crop_overscan_pixels;
	uint32_t crop_left, crop_right, crop_top, crop_bottom;
	uint32_t crop_enable;
	uint32_t downscale_steps;
	uint32_t persist_frame_on_invalid_input;
	uint32_t blend_previous_frame;
	/* A serrated frame is scanned out as its one field at half height instead of being woven; with internal upscaling and
	 * downscale_steps > 0 the halved field is then bob-deinterlaced (video_interface.cpp:1066-1123, 1355-1389). */
	uint32_t upscale_deinterlacing;
	uint32_t vi_aa, vi_scale, vi_serrate, vi_dither_filter, vi_divot_filter, vi_gamma_dither;
} rdp_b200_scanout_options;

/* CommandProcessor::CommandProcessor, rdp_device.cpp:44-162.  `cuda_device` replaces Vulkan::Device&. */
rdp_b200_processor *rdp_b200_create(int cuda_device, void *rdram_ptr, size_t rdram_offset, size_t rdram_size,
                                    size_t hidden_rdram_size, uint32_t flags);
/* ~CommandProcessor, rdp_device.cpp:164-167 */
void rdp_b200_destroy(rdp_b200_processor *p);
/* device_is_supported, rdp_device.cpp:225-228 */
int rdp_b200_device_is_supported(const rdp_b200_processor *p);

/* enqueue_command, rdp_device.cpp:874-897 */
void rdp_b200_enqueue_command(rdp_b200_processor *p, unsigned num_words, const uint32_t *words);
/* enqueue_command_direct, rdp_device.cpp:899-959 */
void rdp_b200_enqueue_command_direct(rdp_b200_processor *p, unsigned num_words, const uint32_t *words);
/* Extension: many back-to-back commands in one call (dump replay / bulk feed). */
void rdp_b200_enqueue_command_stream(rdp_b200_processor *p, size_t num_words, const uint32_t *words);
/* Extension: zero-copy variant.  `words` is borrowed: it must stay valid and unmodified until a later idle(),
 * wait_for_timeline() or scan-out fence covering these commands has returned.  Pinned memory makes the GPU upload
 * asynchronous. */
void rdp_b200_enqueue_command_stream_borrowed(rdp_b200_processor *p, size_t num_words, const uint32_t *words);

/* flush / signal_timeline / wait_for_timeline / idle / begin_frame_context, rdp_device.cpp:169-174,1038-1080 */
void rdp_b200_flush(rdp_b200_processor *p);
uint64_t rdp_b200_signal_timeline(rdp_b200_processor *p);
void rdp_b200_wait_for_timeline(rdp_b200_processor *p, uint64_t index);
void rdp_b200_idle(rdp_b200_processor *p);
void rdp_b200_begin_frame_context(rdp_b200_processor *p);

/* set_validation_interface, rdp_device.cpp:230-233; ValidationInterface::report_rdp_crash and ValidationError,
 * rdp_renderer.hpp:75-95: conditions under which real hardware locks up.  The callback replaces the C++ interface
 * object; like it, it may be called from the command thread.  NULL removes it.  Set it while the processor is idle. */
enum rdp_b200_validation_error
{
	RDP_B200_VALIDATION_FILL_4BPP = 0,
	RDP_B200_VALIDATION_FILL_DEPTH_TEST,
	RDP_B200_VALIDATION_FILL_DEPTH_WRITE,
	RDP_B200_VALIDATION_FILL_IMAGE_READ_ENABLE,
	RDP_B200_VALIDATION_COPY_32BPP,
	RDP_B200_VALIDATION_LOAD_TILE_4BPP,
	RDP_B200_VALIDATION_INVALID_MULTILINE_LOAD_TLUT
};
typedef void (*rdp_b200_validation_callback)(void *user, int error, const char *message);
void rdp_b200_set_validation_callback(rdp_b200_processor *p, rdp_b200_validation_callback callback, void *user);

/* set_quirks, rdp_device.cpp:961-968 */
void rdp_b200_set_quirks(rdp_b200_processor *p, int native_resolution_tex_rect, int native_texture_lod);

/* Memory interaction, rdp_device.cpp:997-1036 */
void *rdp_b200_begin_read_rdram(rdp_b200_processor *p);
void rdp_b200_end_write_rdram(rdp_b200_processor *p);
/* RDRAM coherency (the CUDA replacement for VK_EXT_external_memory_host; the reference's incoherent path,
 * rdp_renderer.cpp:2523-2907).
 * Default: `rdram_ptr` behaves like the reference's imported host memory.  The CPU writes it without telling anyone;
 * whatever a command is about to make the GPU read -- a texture load's source, the colour / depth image of a render
 * pass, the scan-out's framebuffer -- is uploaded first, except where a pass has rendered and the result has not come
 * back yet ("GPU wins", as in the reference), and every pass's output is copied back right behind it.  That is two
 * framebuffers' worth of PCIe traffic per pass: the reference's cost when it cannot import host memory.
 * RDP_B200_FLAG_EXPLICIT_RDRAM_SYNC: the caller declares CPU writes (rdp_b200_mark_rdram_written, or
 * rdp_b200_end_write_rdram for "anything") and asks for rendered memory with rdp_b200_signal_timeline(); nothing else
 * crosses the bus.  The declarations work in the default mode as well.
 *
 * rdp_b200_mark_rdram_written: the CPU wrote RDRAM bytes [offset, offset + size).  May be called at any
 * time from the emulator thread.  The range is uploaded -- whole 1 KiB pages inside it, exact bytes at its two ends, so
 * nothing the CPU did not declare ever overwrites what the GPU rendered next to it -- before the next command that
 * makes the GPU read RDRAM (primitive, texture load, scan-out), in command order.  rdp_b200_end_write_rdram() is the
 * same statement about the whole of RDRAM.  In the other direction every timeline signal copies back exactly the byte
 * ranges passes rendered to since the previous one (asynchronously; rdp_b200_wait_for_timeline => they have landed).
 * CPU writes that are never declared are not seen by the GPU: see INTEGRATION.md "RDRAM coherency". */
void rdp_b200_mark_rdram_written(rdp_b200_processor *p, size_t offset, size_t size);
/* out[0..3] = bytes / copies host -> device, bytes / copies device -> host since creation. */
void rdp_b200_rdram_mirror_stats(rdp_b200_processor *p, uint64_t out4[4]);
void *rdp_b200_begin_read_hidden_rdram(rdp_b200_processor *p);
void rdp_b200_end_write_hidden_rdram(rdp_b200_processor *p);
size_t rdp_b200_get_rdram_size(const rdp_b200_processor *p);
size_t rdp_b200_get_hidden_rdram_size(const rdp_b200_processor *p);
void *rdp_b200_get_tmem(rdp_b200_processor *p);

/* set_vi_register, rdp_device.cpp:970-975 */
void rdp_b200_set_vi_register(rdp_b200_processor *p, uint32_t reg, uint32_t value);
/* Extension: all 14 VI registers in one call, in register-number order (what a dump replayer or an emulator's
 * vertical-interrupt handler has at hand); equivalent to 14 rdp_b200_set_vi_register calls. */
void rdp_b200_set_vi_registers(rdp_b200_processor *p, const uint32_t values[14]);
/* Per-scanline VI registers, rdp_device.cpp:977-995 */
void rdp_b200_begin_vi_register_per_scanline(rdp_b200_processor *p, uint32_t flags);
void rdp_b200_set_vi_register_for_scanline(rdp_b200_processor *p, uint32_t reg_bit, uint32_t value);
void rdp_b200_latch_vi_register_for_scanline(rdp_b200_processor *p, unsigned vi_line);
void rdp_b200_end_vi_register_per_scanline(rdp_b200_processor *p);

/* scanout, rdp_device.cpp:1082-1115: returns a device pointer to RGBA8 pixels (NULL when blank). */
const void *rdp_b200_scanout(rdp_b200_processor *p, const rdp_b200_scanout_options *opts, unsigned *width, unsigned *height);
/* scanout_sync, rdp_device.cpp:1160-1182: copies up to `capacity_pixels` RGBA8 pixels into `rgba`. */
void rdp_b200_scanout_sync(rdp_b200_processor *p, const rdp_b200_scanout_options *opts, uint32_t *rgba,
                           size_t capacity_pixels, unsigned *width, unsigned *height);
/* scanout_async_buffer, rdp_device.cpp:1130-1158 ({Buffer, Fence} pair).  Snapshots the VI registers and queues a
 * scan-out + asynchronous copy into `rgba` (pinned memory recommended) behind the commands enqueued so far; returns at
 * once with a fence for rdp_b200_scanout_wait.  *width / *height (0 x 0: blank or buffer too small) and the pixels
 * are valid once the wait has returned; `rgba`, `width` and `height` must stay alive until then. */
uint64_t rdp_b200_scanout_async(rdp_b200_processor *p, const rdp_b200_scanout_options *opts, uint32_t *rgba,
                                size_t capacity_pixels, unsigned *width, unsigned *height);
void rdp_b200_scanout_wait(rdp_b200_processor *p, uint64_t fence);
void rdp_b200_default_scanout_options(rdp_b200_scanout_options *opts);

/* Introspection (not in the reference API): backend name, kernels launched, primitives / passes submitted. */
const char *rdp_b200_backend_name(void);
uint64_t rdp_b200_launch_count(rdp_b200_processor *p);
uint64_t rdp_b200_primitive_count(rdp_b200_processor *p);
uint64_t rdp_b200_pass_count(rdp_b200_processor *p);
void *rdp_b200_cuda_stream(rdp_b200_processor *p);
/* Name of the tile-kernel instantiation the most recent render pass ran ("generic" = the uber-kernel fallback). */
const char *rdp_b200_last_tile_variant(rdp_b200_processor *p);
/* Seconds the command thread spent: [0] executing ring records, [1] of that submitting passes, [2] of that in queued
 * scan-outs, [3] of that blocked on a staging slot the GPU had not released yet (back-pressure: the GPU is the
 * bottleneck); [0] - [1] - [2] - [3] is command parsing / pass recording. */
void rdp_b200_host_timing(rdp_b200_processor *p, double out4[4]);
/* The backend device handle (rdpb_device*, include/rdp_b200_backend.h) for the measurement entry points. */
void *rdp_b200_backend_device(rdp_b200_processor *p);
/* Extension (stream farms: many processors per host): pins this processor's command thread to the given CPUs, e.g. the
 * cores of the NUMA node its GPU hangs off.  Without it the thread inherits the creating thread's affinity mask (so a
 * process pinned by its launcher needs nothing).  Returns 0, -1 when the platform cannot pin threads. */
int rdp_b200_set_command_thread_affinity(rdp_b200_processor *p, const int *cpus, unsigned num_cpus);

/* ---- Internal upscaling (RDP_B200_FLAG_UPSCALING_2X/4X/8X; ref rdp_device.hpp:43-51, rdp_renderer.cpp:545-600) -------
 * Eligible passes are rendered a second time at S x S the resolution into a separate upscaled domain of S*S RDRAM
 * "slices"; native RDRAM keeps the bit-exact native-resolution result.  Upscaled pixel (x, y) of a framebuffer is
 * native pixel (x / S, y / S) of slice (y % S) * S + x % S (ref shaders/memory_interfacing.h:286-302). */
unsigned rdp_b200_upscaling_factor(rdp_b200_processor *p); /* 1 when off */
uint64_t rdp_b200_upscaled_pass_count(rdp_b200_processor *p);
/* Idles, then copies the upscaled domain to the host: S*S images of get_rdram_size() bytes into rdram_dst and S*S
 * images of get_rdram_size() / 2 bytes into hidden_dst (either may be NULL).  Returns 0, -1 when upscaling is off. */
int rdp_b200_read_upscaled_rdram(rdp_b200_processor *p, void *rdram_dst, void *hidden_dst);
/* RDP_B200_FLAG_SUPER_SAMPLED_READ_BACK (with an upscaling flag; refused without one, ref rdp_renderer.cpp:60-61):
 * passes are rendered in the upscaled domain only, and every native pixel the pass stored to takes the rounded average
 * of its S*S samples (16 / 32-bit colour; the first sample for 8-bit colour and for depth) in native RDRAM, hidden
 * RDRAM left alone (ref rdp_renderer.cpp:2296-2334, update_upscaled_domain_resolve.comp).
 * RDP_B200_FLAG_SUPER_SAMPLED_DITHER adds the reference's 4x4 ordered dither when averaging 5/5/5/1 colours. */

/* ---- RDPDUMP2 command-stream dumps (replaces the reference's rdp_dump.cpp:29-185 DumpPlayer and
 * parallel-rdp/rdp_dump_write.cpp:34-151 RDPDumpWriter; same file format, files are interchangeable) ------------- */
typedef struct rdp_b200_dump_player rdp_b200_dump_player;
typedef struct rdp_b200_dump_writer rdp_b200_dump_writer;

enum
{
	RDP_B200_DUMP_RDP_COMMAND = 2,         /* arg0 = op, words / num_words */
	RDP_B200_DUMP_SET_VI_REGISTER = 3,     /* arg0 = register index, arg1 = value */
	RDP_B200_DUMP_END_FRAME = 4,
	RDP_B200_DUMP_SIGNAL_COMPLETE = 5,
	RDP_B200_DUMP_EOF = 6,
	RDP_B200_DUMP_UPDATE_RDRAM = 7,        /* data / size = the whole current RDRAM image */
	RDP_B200_DUMP_UPDATE_HIDDEN_RDRAM = 9  /* data / size = the whole current hidden RDRAM image */
};

typedef struct rdp_b200_dump_event
{
	uint32_t kind, arg0, arg1, num_words;
	const uint32_t *words; /* valid until the next call on the same player */
	const uint8_t *data;
	size_t size;
} rdp_b200_dump_event;

/* NULL when the file is missing or not an RDPDUMP2 dump of 4/8 MiB RDRAM + 4 MiB hidden RDRAM (ref: rdp_dump.cpp:60-63). */
rdp_b200_dump_player *rdp_b200_dump_open(const char *path);
void rdp_b200_dump_close(rdp_b200_dump_player *d);
size_t rdp_b200_dump_rdram_size(const rdp_b200_dump_player *d);
size_t rdp_b200_dump_hidden_rdram_size(const rdp_b200_dump_player *d);
int rdp_b200_dump_rewind(rdp_b200_dump_player *d);
/* Next record; 0 at the end of the file or on a malformed record (ref: DumpPlayer::iterate). */
int rdp_b200_dump_next(rdp_b200_dump_player *d, rdp_b200_dump_event *ev);
/* Feed records into a processor (the reference's ParallelReplayer, replayer_driver_parallel.cpp:113-160) until an
 * END_FRAME, SIGNAL_COMPLETE or EOF record, which is returned in *stop (SIGNAL_COMPLETE flushes first).  Returns the
 * number of RDP commands enqueued, -1 on a malformed file. */
long rdp_b200_dump_replay(rdp_b200_dump_player *d, rdp_b200_processor *p, rdp_b200_dump_event *stop);

rdp_b200_dump_writer *rdp_b200_dump_writer_open(const char *path, uint32_t rdram_size, uint32_t hidden_rdram_size);
void rdp_b200_dump_writer_flush_dram(rdp_b200_dump_writer *w, const void *rdram, uint32_t size);
void rdp_b200_dump_writer_flush_hidden_dram(rdp_b200_dump_writer *w, const void *hidden, uint32_t size);
void rdp_b200_dump_writer_command(rdp_b200_dump_writer *w, uint32_t op, const uint32_t *words, uint32_t num_words);
void rdp_b200_dump_writer_set_vi_register(rdp_b200_dump_writer *w, uint32_t reg, uint32_t value);
void rdp_b200_dump_writer_end_frame(rdp_b200_dump_writer *w);
void rdp_b200_dump_writer_signal_complete(rdp_b200_dump_writer *w);
void rdp_b200_dump_writer_close(rdp_b200_dump_writer *w); /* writes EOF */

#ifdef __cplusplus
}
#endif
#endif
