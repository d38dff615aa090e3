// rdp_renderer.hpp -- render-pass recording and submission (host half of the hot path).
//
// Role mirrors parallel-rdp/rdp_renderer.{hpp,cpp}: tracks RDP state, batches primitives into render passes,
// applies the semantic flush rules, decomposes texture loads and sequences the GPU stages.  What differs by
// design: primitives are forwarded as raw command words (decode runs on the GPU), state is snapshotted by a
// dirty flag instead of memcmp-searched caches, and a pass holds up to 16 Ki primitives.
#pragma once
#include "rdp_b200_backend.h"
#include "rdram_mirror.hpp"
#include "rdp_types.h"
#include "rdp_spec.h"
#include <stddef.h>
#include <chrono>
#include <stdint.h>
#include <vector>

namespace RDP
{
enum class FBFormat : uint32_t
{
	I4 = 0,
	I8 = 1,
	RGBA5551 = 2,
	IA88 = 3,
	RGBA8888 = 4
};

enum class TextureSize : uint8_t
{
	Bpp4 = 0,
	Bpp8 = 1,
	Bpp16 = 2,
	Bpp32 = 3
};

enum class TextureFormat : uint8_t
{
	RGBA = 0,
	YUV = 1,
	CI = 2,
	IA = 3,
	I = 4
};

enum class UploadMode : uint32_t
{
	Tile = 0,
	TLUT = 1,
	Block = 2
};

struct LoadTileInfo
{
	uint32_t tex_addr;
	uint32_t tex_width;
	TextureFormat fmt;
	TextureSize size;
	uint32_t slo, tlo, shi, thi;
	UploadMode mode;
};

// Conditions under which real hardware locks up (reference: rdp_renderer.hpp:75-95).
enum class ValidationError
{
	Fill4bpp,
	FillDepthTest,
	FillDepthWrite,
	FillImageReadEnable,
	Copy32bpp,
	LoadTile4bpp,
	InvalidMultilineLoadTlut
};

struct ValidationInterface
{
	virtual ~ValidationInterface() = default;
	// May be called from the command thread.
	virtual void report_rdp_crash(ValidationError err, const char *msg) = 0;
};

// What the renderer tells the owner of the RDRAM mirror around every pass (implicit coherency; rdp_device.hpp).
struct RdramCoherencyInterface
{
	virtual ~RdramCoherencyInterface() = default;
	// The pass about to be submitted reads RDRAM bytes [addr, addr + bytes) (colour / depth image).
	virtual void gpu_will_read(uint32_t addr, uint32_t bytes) = 0;
	// A pass that renders has been submitted (its written ranges have been noted in the mirror).
	virtual void pass_submitted() = 0;
};

class Renderer
{
public:
	bool init(rdpb_device *device, size_t rdram_size);
	// Internal upscaling factor 1 << log2 (0: off).  Allocates the upscaled domain in the backend.
	bool set_upscaling(uint32_t log2);
	// Super-sampled read-back (ref RendererOptions, rdp_renderer.hpp:71-72): passes are rendered upscaled only and
	// native RDRAM takes the average of each pixel's samples.  Needs upscaling; returns false otherwise.
	bool set_super_sampled_readback(bool enable, bool dither);
	// QuirkFlags (csrc/rdp_types.h); part of every state snapshot taken from now on.
	void set_quirks(uint32_t quirk_flags);
	uint32_t get_upscaling_log2() const { return upscale_log2; }
	uint64_t get_upscaled_pass_count() const { return upscaled_pass_counter; }
	void set_validation_interface(ValidationInterface *iface) { validation = iface; }

	// Register state.  Callers mutate `regs` then call mark_state_dirty().
	rdpb::PassState &registers() { return regs; }
	void mark_state_dirty() { state_dirty = true; }
	void set_other_modes(uint32_t w0, uint32_t w1);
	void set_combine(uint32_t w0, uint32_t w1);

	void set_color_framebuffer(uint32_t addr, uint32_t width, FBFormat fmt);
	void set_depth_framebuffer(uint32_t addr);

	void draw_primitive(const uint32_t *words, unsigned num_words);
	// Consumes the leading run of triangle commands of `words`; returns the words consumed (rdp_renderer.cpp).
	size_t draw_triangle_run(const uint32_t *words, size_t num_words);
	uint32_t num_words_in_pass() const { return num_words; }
	// While a borrowed source is active, primitives whose words lie inside it are uploaded from there (no host copy).
	void begin_borrowed_source(const uint32_t *words, size_t num_words) { borrowed_base = words; borrowed_end = words + num_words; }
	void end_borrowed_source() { borrowed_base = borrowed_end = nullptr; }
	void load_tile(uint32_t tile, const LoadTileInfo &info);

	void flush_queues();
	// Where passes note the byte ranges they render to (device -> host side of the RDRAM mirror).
	void set_rdram_mirror(RdramMirror *m) { mirror = m; }
	void set_rdram_coherency(RdramCoherencyInterface *c) { coherency = c; }
	void invalidate_tmem_snapshot() { tmem_snapshot_valid = false; }

	uint64_t get_primitive_counter() const { return base_primitive_index; }
	uint64_t get_pass_counter() const { return pass_counter; }
	uint64_t get_flush_ns() const { return flush_ns; }
	uint64_t get_backpressure_ns() const { return backpressure_ns; }

private:
	rdpb_device *device = nullptr;
	ValidationInterface *validation = nullptr;
	size_t rdram_size = 0;

	rdpb::PassState regs = {};
	bool state_dirty = true;
	uint32_t other_modes[2] = {};

	struct
	{
		uint32_t addr = 0, depth_addr = 0, width = 0;
		FBFormat fmt = FBFormat::I4;
		uint32_t deduced_height = 0;
		bool color_write_pending = false, depth_write_pending = false;
	} fb;

	// Pass being recorded.
	rdpb_pass_buffers pass = {};
	bool pass_open = false;
	uint32_t num_words = 0, num_prims = 0, num_states = 0, num_uploads = 0, num_spans = 0, num_jobs = 0;
	// State snapshots are interned per pass (ref StateCache, rdp_data_structures.hpp:256-362): a stream that toggles between
	// a few states does not grow the table.  Open addressing over the snapshot's hash; entries are index + 1.
	enum { StateTableSize = 4096 };
	uint16_t state_table[StateTableSize] = {};
	uint64_t state_hash[rdpb::kMaxStatesPerPass] = {};
	uint32_t current_state = 0;
	bool pass_reads_tmem = false;
	// Where the pass's command words come from: runs of host words and their offsets in the device word buffer.
	// Staged runs point into pass.words; borrowed runs point into the caller's stream (gaps between primitives --
	// state commands -- ride along, which is cheaper than cutting the run).
	struct WordSegment
	{
		const uint32_t *src;
		uint32_t count, dst;
		bool borrowed;
	};
	std::vector<WordSegment> segments;
	uint32_t staged_words = 0;
	const uint32_t *borrowed_base = nullptr, *borrowed_end = nullptr;
	uint32_t place_words(const uint32_t *words, unsigned count);
	// What the pass's non-fill states have in common; lets the backend pick a specialised tile kernel.
	rdpb::PassSummary summary = {};
	uint32_t state_tile_kind_bits[8] = {}; // texel kind bit of each tile descriptor under the current snapshot
	void reset_summary();
	uint32_t base_primitive_index = 0;
	uint64_t pass_counter = 0;
	uint64_t upscaled_pass_counter = 0;
	// Byte ranges [lo, hi) the previous pass renders to (colour, depth), and whether a texture load of the pass being
	// recorded may read from them (rdpb_launch_front's ordering hint).
	uint64_t rendered_ranges[2][2] = {};
	bool rendered_ranges_valid = false;
	bool uploads_may_read_previous_pass = false;
	uint32_t upscale_log2 = 0;
	bool super_sampled_readback = false, super_sampled_dither = false;
	uint64_t flush_ns = 0; // time spent submitting passes (written by the thread that drives the renderer)
	uint64_t backpressure_ns = 0; // time blocked waiting for a staging slot (the GPU is behind)
	bool tmem_snapshot_valid = false;

	RdramMirror *mirror = nullptr;
	RdramCoherencyInterface *coherency = nullptr;

	void ensure_pass();
	void snapshot_state();
	void finalize_state(rdpb::PassState &st) const;
	void validate_draw_state() const;
	bool tmem_upload_needs_flush(uint32_t addr) const;
	void load_tile_iteration(uint32_t tile, const LoadTileInfo &info, uint32_t tmem_offset);
	void note_gpu_write(uint32_t addr, uint32_t bytes);
};
}
