// rdp_device.hpp -- RDP::CommandProcessor, the drop-in boundary.
//
// Same public surface as the reference class (parallel-rdp/rdp_device.hpp:111-197): same method names,
// argument meaning, ownership and error behaviour.  The two Vulkan-typed spots change:
//   * the constructor takes a CUDA device selection (RDP::DeviceConfig) instead of Vulkan::Device&,
//   * scanout() returns RDP::ScanoutImage {device pointer, width, height} instead of Vulkan::ImageHandle,
//     and VIScanoutBuffer holds a host vector instead of Vulkan::Buffer + Fence.
// rdram_ptr is borrowed (caller frees after destroying the processor); hidden RDRAM and TMEM are owned here.
// No exceptions, no error codes: failure => device_is_supported() == false plus a line on stderr.
#pragma once
#include "command_ring.hpp"
#include "rdp_renderer.hpp"
#include "video_interface.hpp"
#include "rdp_dump.hpp"
#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

namespace RDP
{
struct RGBA
{
	uint8_t r, g, b, a;
};

enum CommandProcessorFlagBits
{
	COMMAND_PROCESSOR_FLAG_HOST_VISIBLE_HIDDEN_RDRAM_BIT = 1 << 0,
	COMMAND_PROCESSOR_FLAG_HOST_VISIBLE_TMEM_BIT = 1 << 1,
	COMMAND_PROCESSOR_FLAG_UPSCALING_2X_BIT = 1 << 2,
	COMMAND_PROCESSOR_FLAG_UPSCALING_4X_BIT = 1 << 3,
	COMMAND_PROCESSOR_FLAG_UPSCALING_8X_BIT = 1 << 4,
	COMMAND_PROCESSOR_FLAG_SUPER_SAMPLED_READ_BACK_BIT = 1 << 5,
	COMMAND_PROCESSOR_FLAG_SUPER_SAMPLED_DITHER_BIT = 1 << 6,
	// Extension (not a reference flag).  By default the caller's RDRAM behaves like the reference's imported host memory:
	// the CPU writes it without telling anyone and the processor makes the GPU see it -- every range a command is about
	// to read (texture-load sources, the colour / depth image of a pass, the scan-out's framebuffer) is uploaded first
	// unless a pass has rendered to it and the result has not been returned yet, and every pass's output is copied
	// back right after it (the reference's incoherent path: rdp_renderer.cpp:2523-2907).  With this bit the caller
	// promises to declare CPU writes (mark_rdram_written / end_write_rdram) and to ask for the rendered memory with
	// signal_timeline(): nothing else crosses the bus.
	COMMAND_PROCESSOR_FLAG_EXPLICIT_RDRAM_SYNC_BIT = 1 << 16
};
using CommandProcessorFlags = uint32_t;

struct Quirks
{
	Quirks()
	{
		u.options.native_resolution_tex_rect = true;
		u.options.native_texture_lod = false;
	}
	void set_native_resolution_tex_rect(bool enable) { u.options.native_resolution_tex_rect = enable; }
	void set_native_texture_lod(bool enable) { u.options.native_texture_lod = enable; }

	union
	{
		struct Opts
		{
			bool native_resolution_tex_rect;
			bool native_texture_lod;
		} options;
		uint32_t words[1];
	} u;
};

// Replaces Vulkan::Device& in the constructor.
struct DeviceConfig
{
	int cuda_device = 0;
};

class CommandProcessor : private RdramCoherencyInterface
{
public:
	CommandProcessor(const DeviceConfig &device, void *rdram_ptr, size_t rdram_offset, size_t rdram_size,
	                 size_t hidden_rdram_size, CommandProcessorFlags flags);
	~CommandProcessor();
	CommandProcessor(const CommandProcessor &) = delete;
	void operator=(const CommandProcessor &) = delete;

	void set_validation_interface(ValidationInterface *iface);
	bool device_is_supported() const;

	// Synchronization.
	void flush();
	uint64_t signal_timeline();
	void wait_for_timeline(uint64_t index);
	void idle();
	void begin_frame_context();

	// Queues up state and drawing commands.
	void enqueue_command(unsigned num_words, const uint32_t *words);
	void enqueue_command_direct(unsigned num_words, const uint32_t *words);
	// Extension: a packed stream of back-to-back commands (dump replay, benchmarking).
	void enqueue_command_stream(size_t num_words, const uint32_t *words);
	// Extension: same, but the words are *borrowed* -- nothing is copied on the host; primitives are DMA'd to the
	// GPU straight from `words` (pinned memory makes that asynchronous).  The buffer must stay valid and unmodified
	// until a later idle(), wait_for_timeline() or scan-out fence covering these commands has returned.
	void enqueue_command_stream_borrowed(size_t num_words, const uint32_t *words);
	// Called by the command thread for a borrowed record.
	void execute_borrowed_stream(size_t num_words, const uint32_t *words);
	void execute_command_run(size_t num_words, const uint32_t *words);

	void set_quirks(const Quirks &quirks);

	// Interact with memory.  Pointers are valid while the processor is idle.
	void *begin_read_rdram();
	void end_write_rdram();
	// Extension: the CPU wrote [offset, offset + size) of RDRAM (may be called at any time, from the emulator thread).
	// Uploaded -- whole 1 KiB pages inside the range, exact bytes at its ends -- before the next command that makes the
	// GPU read RDRAM.  end_write_rdram() is the same statement about the whole range.
	void mark_rdram_written(size_t offset, size_t size);
	void get_rdram_mirror_stats(uint64_t out[4]) const; // bytes / copies host -> device, bytes / copies device -> host
	void *begin_read_hidden_rdram();
	void end_write_hidden_rdram();
	size_t get_rdram_size() const;
	size_t get_hidden_rdram_size() const;
	void *get_tmem();

	void set_vi_register(VIRegister reg, uint32_t value);

	ScanoutImage scanout(const ScanoutOptions &opts = {});
	void scanout_sync(std::vector<RGBA> &colors, unsigned &width, unsigned &height, const ScanoutOptions &opts = {});
	void scanout_async_buffer(VIScanoutBuffer &buffer, const ScanoutOptions &opts = {});
	// Scan out and start an asynchronous copy of the RGBA8 image into caller memory (pinned memory makes it a true
	// DMA); returns a fence for scanout_wait().  This is scanout_async_buffer's {buffer, fence} pair (reference
	// rdp_device.cpp:1130-1158) with a caller-owned buffer.
	// The call only snapshots the VI registers and queues the scan-out behind the commands enqueued so far: the
	// command thread flushes, scans out, starts the copy and records the fence, so the caller is free to feed the
	// next frame immediately.  *width / *height are written before the fence signals.
	uint64_t scanout_async(void *rgba, size_t capacity_pixels, unsigned *width, unsigned *height, const ScanoutOptions &opts = {});
	void scanout_wait(uint64_t fence);

	void begin_vi_register_per_scanline(VideoInterface::PerScanlineRegisterFlags flags);
	void set_vi_register_for_scanline(VideoInterface::PerScanlineRegisterBits reg, uint32_t value);
	void latch_vi_register_for_scanline(unsigned vi_line);
	void end_vi_register_per_scanline();

	// Introspection for benchmarks / tests (not part of the reference surface).
	rdpb_device *get_backend_device() { return device; }
	uint64_t get_primitive_counter() const { return renderer.get_primitive_counter(); }
	uint64_t get_pass_counter() const { return renderer.get_pass_counter(); }
	uint32_t get_upscaling_log2() const { return renderer.get_upscaling_log2(); }
	uint64_t get_upscaled_pass_count() const { return renderer.get_upscaled_pass_count(); }
	// Seconds the command thread spent [0] executing ring records in total, [1] inside pass submission
	// (Renderer::flush_queues), [2] inside queued scan-outs, [3] blocked on a staging slot the GPU has not released yet
	// (back-pressure).  [0] - [1] - [2] - [3] is command parsing / pass recording.
	void get_host_timing(double out[4]) const;
	void set_command_thread_affinity(const std::vector<int> &cpus) { ring.set_affinity(cpus); }
	void add_ring_busy_ns(uint64_t ns) { ring_busy_ns.fetch_add(ns, std::memory_order_relaxed); }

private:
	rdpb_device *device = nullptr;
	uint8_t *host_rdram = nullptr;      // borrowed (or owned_rdram)
	std::unique_ptr<uint8_t[]> owned_rdram;
	bool host_rdram_registered = false;
	std::vector<uint8_t> host_hidden;
	size_t rdram_size = 0;
	size_t hidden_size = 0;
	CommandProcessorFlags flags = 0;
	uint64_t scanout_fence_counter = 0;
	std::atomic<uint64_t> ring_busy_ns{0}, scanout_ns{0};
	std::atomic<uint64_t> recorded_scanout_fence{0};
	struct ScanoutRequest;
	void execute_scanout_request(ScanoutRequest *req);
	bool is_supported = false;
	bool single_threaded_processing = true;
	// PARALLEL_RDP_DUMP_PATH: everything fed to the processor is also written to an RDPDUMP2 file
	// (ref rdp_device.cpp:60-76, 874-897, 970-975, 1087-1093).
	std::unique_ptr<RDPDumpWriter> dump_writer;
	bool dump_in_command_list = false;
	void dump_begin_command_list();
	void dump_end_frame();

	RdramMirror mirror; // what the CPU wrote since the last upload / what passes rendered since the last read-back
	bool implicit_coherency = true; // !COMMAND_PROCESSOR_FLAG_EXPLICIT_RDRAM_SYNC_BIT
	uint64_t internal_fence_counter = 0;
	// RdramCoherencyInterface (rdp_renderer.hpp): what the renderer tells the processor around every pass.
	void gpu_will_read(uint32_t addr, uint32_t bytes) override;
	void pass_submitted() override;
	std::mutex idle_lock;

	Renderer renderer;
	VideoInterface vi;
	CommandRing ring;

	struct
	{
		uint32_t addr = 0, width = 0;
		TextureFormat fmt = TextureFormat::RGBA;
		TextureSize size = TextureSize::Bpp4;
	} texture_image;

	uint64_t timeline_value = 0;
	std::atomic<uint64_t> signalled_timeline{0};
	Quirks quirks;
	void apply_quirks();

	void enqueue_command_inner(unsigned num_words, const uint32_t *words);
	void enqueue_meta(RingMeta kind, unsigned num_words, const uint32_t *payload);
	void execute_meta(RingMeta kind, const uint32_t *payload, unsigned num_words);
	void drain_command_ring();
	void sync_host_rdram_to_device();
	bool readback_gpu_writes(uint64_t fence_value, bool wait);
	void flush_and_signal();

	void op_set_color_image(const uint32_t *w);
	void op_set_mask_image(const uint32_t *w);
	void op_set_scissor(const uint32_t *w);
	void op_set_texture_image(const uint32_t *w);
	void op_set_tile(const uint32_t *w);
	void op_set_tile_size(const uint32_t *w);
	void op_load(const uint32_t *w, UploadMode mode);
	void op_set_convert(const uint32_t *w);
	void op_set_key_gb(const uint32_t *w);
	void op_set_key_r(const uint32_t *w);
	void op_set_prim_color(const uint32_t *w);
	void op_set_prim_depth(const uint32_t *w);

	friend class CommandRing;
};
}
