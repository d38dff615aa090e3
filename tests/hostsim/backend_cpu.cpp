// tests/hostsim/backend_cpu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A CPU build of the *same* device functions the CUDA kernels execute (parallel-rdp_b200/csrc/*.cuh compiled
// with g++), behind the same backend C-ABI (include/rdp_b200_backend.h).  It exists so the host logic and the
// integer pixel pipeline can be diffed against the Angrylion oracle in a container without a GPU
// (`pytest -m "not gpu"`).  It is never linked into librdp_b200.so and is not a fallback: the product
// library fails loudly when CUDA is unavailable.
#include "rdp_b200_backend.h"
#include "rdp_pass.cuh"
#include "rdp_vi.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

using namespace rdpb;

struct rdpb_device
{
	uint32_t rdram_size = 0, hidden_size = 0;
	std::vector<uint8_t> rdram, hidden, tmem, tmem_instances, blend_lut, gamma_lut;
	std::vector<uint32_t> words, staging_words;
	std::vector<PrimRecord> records;
	std::vector<PassState> states;
	std::vector<TmemUpload> uploads;
	std::vector<PrimSetup> prims;
	std::vector<Span> spans;
	std::vector<uint32_t> fine, coarse;
	// internal upscaling
	uint32_t scale_log2 = 0;
	std::vector<uint8_t> up_rdram, up_hidden, up_reference;
	std::vector<Span> up_spans;
	std::vector<uint32_t> up_fine, up_coarse;
	rdpb_pass_counts counts = {};
	std::vector<uint32_t> vi_stage[3];
	std::vector<uint32_t> vi_out, vi_prev;
	const uint32_t *vi_result = nullptr; // what the last scan-out produced (vi_out, or a down-sized copy)
	uint64_t launches = 0;
	uint64_t deferred_passes = 0;
	bool last_deferred = false;
	unsigned long long counters[8] = {};
	// RDP_HOSTSIM_HIZ=1 (test-only experiment, see hiz_check_tile): entries seen / rejected by the bound with the entry's Z
	// bounded from its spans / with no passing fragment at all / rejected although a fragment passed (must stay 0) /
	// rejected when the entry's smallest fragment Z is known exactly / span bound above a real fragment's Z (must stay 0) /
	// rejected with per-row instead of per-tile maxima (exact per-row Z) / of those, wrongly (must stay 0)
	unsigned long long hiz[8] = {};
	// same switch: (primitive, tile) entries a pass would have with other tile shapes, and the pixels inside the spans
	// ([0] 8x4, [1] 16x4, [2] 8x8, [3] 16x8, [4] covered pixels)
	unsigned long long tile_shapes[5] = {};
	// same switch: fragments entering phase A (coverage + Z) / surviving phase B (what phases C and D shade and blend) /
	// fragments of entries the span-based bound rejects (phase A + B work it would save)
	unsigned long long hiz_fragments[3] = {};
	std::string error;

	PassBuffers buffers()
	{
		PassBuffers b;
		b.words = words.data();
		b.records = records.data();
		b.states = states.data();
		b.uploads = uploads.data();
		b.prims = prims.data();
		b.spans = spans.data();
		b.fine_mask = fine.data();
		b.coarse_mask = coarse.data();
		b.jobs = nullptr;
		b.job_counter = nullptr;
		b.tmem = tmem.data();
		b.tmem_instances = tmem_instances.data();
		b.rdram = rdram.data();
		b.hidden = hidden.data();
		b.blend_lut = blend_lut.data();
		b.counters = counters;
		return b;
	}

	// Buffers of an upscaled pass: its own RDRAM domain, spans and masks; everything else is the native pass's.
	PassBuffers upscaled_buffers(const FbInfo &fb)
	{
		PassBuffers b = buffers();
		const size_t tiles = size_t(fb.tiles_x) * fb.tiles_y;
		if (up_fine.size() < tiles * fb.mask_stride)
			up_fine.resize(tiles * fb.mask_stride, 0);
		if (up_coarse.size() < tiles * fb.coarse_stride)
			up_coarse.resize(tiles * fb.coarse_stride, 0);
		b.spans = up_spans.data();
		b.fine_mask = up_fine.data();
		b.coarse_mask = up_coarse.data();
		b.rdram = up_rdram.data();
		b.hidden = up_hidden.data();
		return b;
	}
};

// RDP_HOSTSIM_NULL=1 turns every launch into a no-op: used to time the host front-end alone (tools/host_timing.py).
static const bool g_null_compute = getenv("RDP_HOSTSIM_NULL") != nullptr;

extern "C" {

const char *rdpb_backend_name(void) { return "hostsim-cpu (test only)"; }

rdpb_device *rdpb_device_create(int, uint32_t rdram_size, uint32_t hidden_size)
{
	auto *d = new rdpb_device;
	d->rdram_size = rdram_size;
	d->hidden_size = hidden_size;
	d->rdram.assign(rdram_size, 0);
	d->hidden.assign(hidden_size, 0);
	d->tmem.assign(kTmemBytes, 0);
	d->tmem_instances.assign(size_t(kTmemBytes) * (kMaxUploadsPerPass + 1), 0);
	d->blend_lut.resize(0x8000);
	build_blender_lut(d->blend_lut.data());
	d->gamma_lut.resize(256 + 256 * 64);
	build_gamma_lut(d->gamma_lut.data());
	d->words.resize(kMaxPassWords);
	d->staging_words.resize(kMaxPassWords);
	d->records.resize(kMaxPrimsPerPass);
	d->states.resize(kMaxStatesPerPass);
	d->uploads.resize(kMaxUploadsPerPass);
	d->prims.resize(kMaxPrimsPerPass);
	d->spans.resize(kMaxSpansPerPass);
	d->fine.assign(size_t(kMaxTilesX) * kMaxTilesY * (kMaxPrimsPerPass / 32), 0);
	d->coarse.assign(size_t(kMaxTilesX) * kMaxTilesY * (kMaxPrimsPerPass / 1024), 0);
	return d;
}

void rdpb_device_destroy(rdpb_device *dev) { delete dev; }
const char *rdpb_last_error(rdpb_device *dev) { return dev->error.c_str(); }
int rdpb_register_host_memory(rdpb_device *, void *, size_t) { return 0; }
int rdpb_unregister_host_memory(rdpb_device *, void *) { return 0; }

int rdpb_rdram_write(rdpb_device *dev, uint32_t offset, const void *src, uint32_t size, int)
{
	memcpy(dev->rdram.data() + offset, src, size);
	return 0;
}
int rdpb_rdram_read(rdpb_device *dev, uint32_t offset, void *dst, uint32_t size, int)
{
	memcpy(dst, dev->rdram.data() + offset, size);
	return 0;
}
int rdpb_rdram_read_fenced(rdpb_device *dev, void *host_base, const uint32_t *ranges, uint32_t num_ranges, uint64_t)
{
	for (uint32_t i = 0; i < num_ranges; i++)
		memcpy(static_cast<uint8_t *>(host_base) + ranges[2 * i], dev->rdram.data() + ranges[2 * i], ranges[2 * i + 1]);
	return 0;
}
int rdpb_hidden_write(rdpb_device *dev, uint32_t offset, const void *src, uint32_t size, int)
{
	memcpy(dev->hidden.data() + offset, src, size);
	return 0;
}
int rdpb_hidden_read(rdpb_device *dev, uint32_t offset, void *dst, uint32_t size, int)
{
	memcpy(dst, dev->hidden.data() + offset, size);
	return 0;
}
int rdpb_hidden_fill(rdpb_device *dev, uint8_t value)
{
	memset(dev->hidden.data(), value, dev->hidden.size());
	return 0;
}
void *rdpb_tmem_host_ptr(rdpb_device *dev) { return dev->tmem.data(); }
void *rdpb_rdram_device_ptr(rdpb_device *dev) { return dev->rdram.data(); }
void *rdpb_hidden_device_ptr(rdpb_device *dev) { return dev->hidden.data(); }

int rdpb_pass_begin(rdpb_device *dev, rdpb_pass_buffers *out)
{
	out->words = dev->staging_words.data();
	out->records = dev->records.data();
	out->states = dev->states.data();
	out->uploads = dev->uploads.data();
	return 0;
}

int rdpb_pass_upload(rdpb_device *dev, const rdpb_pass_counts *counts)
{
	dev->counts = *counts;
	return 0;
}

int rdpb_pass_upload_words(rdpb_device *dev, const uint32_t *src, uint32_t num_words, uint32_t dst_word_offset)
{
	memcpy(dev->words.data() + dst_word_offset, src, size_t(num_words) * 4);
	return 0;
}

int rdpb_launch_front(rdpb_device *dev, int update_tmem, uint32_t num_uploads, uint32_t num_prims)
{
	update_tmem &= 1; // bit 1 is a stream-ordering hint for the CUDA backend
	if (g_null_compute)
		return 0;
	PassBuffers b = dev->buffers();
	if (update_tmem)
		for (int slot = 0; slot < int(kTmemBytes / 2); slot++)
			pass_tmem_word(b, num_uploads, dev->rdram_size, slot);
	for (uint32_t i = 0; i < num_prims; i++)
		pass_setup_prim(b, i);
	dev->launches++;
	return 0;
}

int rdpb_launch_span_bin(rdpb_device *dev, const void *fb_info)
{
	if (g_null_compute)
		return 0;
	const FbInfo fb = *static_cast<const FbInfo *>(fb_info);
	const int sl = int(fb.scale_log2), S = 1 << sl;
	PassBuffers b = sl ? dev->upscaled_buffers(fb) : dev->buffers();
	auto or_fn = [](uint32_t *p, uint32_t v) { *p |= v; };
	for (uint32_t i = 0; i < fb.num_prims; i++)
	{
		const PrimSetup &p = b.prims[i];
		if (p.flags & SF_NO_PIXELS)
			continue;
		for (int y = p.line_lo * S; y <= (p.line_hi + 1) * S + S - 1; y++)
		{
			int lo, hi;
			pass_span_line(b, fb, p, y, lo, hi, sl);
			if (lo <= hi)
				pass_bin_row(fb, i, y / int(kTileH), lo, hi, or_fn, b.fine_mask, b.coarse_mask);
		}
	}
	dev->launches++;
	return 0;
}

// ---- experiment: conservative per-(tile, primitive) depth rejection ----------------------------------
// Walks one tile primitive-major over private copies of its 32 pixels' depth state, exactly as phases A and B of the
// deferred pipeline would (defer_candidate / defer_depth), and checks the rejection bound derived in DESIGN.md: an entry
// in opaque Z-compare mode whose smallest fragment Z exceeds max(tileA, tileZ + 8 * P(dz)) cannot pass anywhere in the
// tile.  tileA / tileZ are exact at the start of every 32-primitive batch and widened inside the batch by what the
// earlier entries of the batch could have written.  Nothing is stored: the regular walk that follows renders the tile.
static int hiz_pow2_floor(int v) { return v > 0 ? 1 << find_msb(uint32_t(v)) : 0; }

// Lower bound of the Z of every fragment the primitive can have in the tile at (x0, y0), from what the mask phase reads
// anyway: the four lines' span headers.  Along a line Z is linear in x before the snapping, the centroid offsets add
// between 0 and 3 quarter-steps of dz/dx and dz/dy (interpolate_depth), the final shift and clamp are monotone as long as
// nothing wraps; if anything can wrap, no bound is given (-1).
static int hiz_entry_z_lower_bound(const PassBuffers &b, const FbInfo &fb, const PrimSetup &p, int x0, int y0)
{
	long long t_lo = 0, t_hi = 0;
	bool any = false;
	const long long dzdx = p.dstzw_dx[2];
	const long long ox = 3ll * (p.dstzw_dx[2] >> 10), oy = 3ll * (p.dstzw_dy[2] >> 10);
	const long long off_lo = (ox < 0 ? ox : 0) + (oy < 0 ? oy : 0), off_hi = (ox > 0 ? ox : 0) + (oy > 0 ? oy : 0);
	for (int ly = 0; ly < int(kTileH); ly++)
	{
		const int y = y0 + ly;
		if (y >= int(fb.height) || y < p.line_lo || y > p.line_hi)
			continue;
		const Span &sp = b.spans[p.span_base + uint32_t(y - p.line_lo)];
		if (!sp.valid)
			continue;
		const int xlo = imax(int(sp.start_x), x0), xhi = imin(imin(int(sp.end_x), x0 + int(kTileW) - 1), int(fb.width) - 1);
		if (xlo > xhi)
			continue;
		for (int k = 0; k < 2; k++)
		{
			const long long z = (long long)(sp.stzw[2]) + dzdx * ((k ? xhi : xlo) - sp.base_x);
			if (z < -2147483648ll || z > 2147483647ll)
				return -1;
			const long long snapped = (z >> 10) * 4;
			const long long lo = (snapped + off_lo) >> 5, hi = (snapped + off_hi) >> 5;
			t_lo = any ? (lo < t_lo ? lo : t_lo) : lo;
			t_hi = any ? (hi > t_hi ? hi : t_hi) : hi;
			any = true;
		}
	}
	if (!any || t_lo < -131072 || t_hi >= 393216) // clamp_depth's 19-bit window
		return -1;
	return int(t_lo < 0 ? 0 : (t_lo > 0x3ffff ? 0x3ffff : t_lo));
}

static void hiz_check_tile(rdpb_device *dev, const PassBuffers &b, const FbInfo &fb, uint32_t tile, uint32_t tx, uint32_t ty)
{
	constexpr int kInf = 0x7fffffff;
	FbView view[32];
	PixelState px[32];
	bool inside[32];
	for (int l = 0; l < 32; l++)
	{
		const int x = int(tx * kTileW) + (l & 7), y = int(ty * kTileH) + (l >> 3);
		inside[l] = x < int(fb.width) && y < int(fb.height);
		pass_pixel_begin<SpecGeneric>(b, fb, view[l], px[l], x, y, 0);
	}
	const uint32_t *coarse = b.coarse_mask + size_t(tile) * fb.coarse_stride;
	const uint32_t *fine = b.fine_mask + size_t(tile) * fb.mask_stride;
	for (uint32_t group = 0; group < fb.mask_stride; group++)
	{
		if (!((coarse[group >> 5] >> (group & 31u)) & 1u))
			continue;
		// exact bounds at batch start (whole tile, and per row for the per-row variant)
		int cap_a = 0, cap_z = 0;
		int row_a[4] = {}, row_z[4] = {};
		for (int l = 0; l < 32; l++)
		{
			if (!inside[l])
				continue;
			const int mem_depth = px[l].depth, mem_dz_c = px[l].dz;
			const int memory_z = z_decompress(mem_depth);
			int memory_dz = 1 << mem_dz_c;
			const int precision = (mem_depth >> 11) & 0xf;
			bool unbounded = memory_z == 0x3ffff;
			if (precision < 3)
			{
				if (memory_dz != 0x8000)
					memory_dz = imax(memory_dz << 1, 16 >> precision);
				else
					unbounded = true; // coplanar
			}
			cap_z = imax(cap_z, memory_z);
			cap_a = unbounded ? kInf : imax(cap_a, memory_z + 8 * hiz_pow2_floor(memory_dz));
			row_z[l >> 3] = imax(row_z[l >> 3], memory_z);
			row_a[l >> 3] = unbounded ? kInf : imax(row_a[l >> 3], memory_z + 8 * hiz_pow2_floor(memory_dz));
		}
		uint32_t bits = fine[group];
		while (bits)
		{
			const uint32_t index = group * 32u + uint32_t(find_lsb(bits));
			bits &= bits - 1u;
			const PrimSetup &p = b.prims[index];
			const PassState &st = b.states[p.state_id];
			uint32_t cand[32];
			int zmin = kInf, zmax = -1;
			int rmin[4] = { kInf, kInf, kInf, kInf }, rmax[4] = { -1, -1, -1, -1 };
			bool any = false, fill = false;
			for (int l = 0; l < 32; l++)
			{
				cand[l] = 0;
				const int x = int(tx * kTileW) + (l & 7), y = int(ty * kTileH) + (l >> 3);
				if (!inside[l] || y < p.line_lo || y > p.line_hi)
					continue;
				const Span &sp = b.spans[p.span_base + uint32_t(y - p.line_lo)];
				if (!sp.valid)
					continue;
				cand[l] = defer_candidate<SpecGeneric>(st, p, sp, x);
				if (!(cand[l] & DC_VALID))
					continue;
				any = true;
				if (cand[l] & DC_FILL)
					fill = true;
				else
				{
					zmin = imin(zmin, int(cand[l] & DC_Z_MASK));
					zmax = imax(zmax, int(cand[l] & DC_Z_MASK));
					rmin[l >> 3] = imin(rmin[l >> 3], int(cand[l] & DC_Z_MASK));
					rmax[l >> 3] = imax(rmax[l >> 3], int(cand[l] & DC_Z_MASK));
				}
			}
			if (!any || fill)
				continue; // fill-mode entries never touch depth in a deferrable pass
			const bool opaque_compare = (st.blend_flags & BF_Z_COMPARE) != 0u && st.z_mode == 0;
			const int own = 8 * hiz_pow2_floor(int(p.dz));
			const int zlb = hiz_entry_z_lower_bound(b, fb, p, int(tx * kTileW), int(ty * kTileH));
			const bool ideal = opaque_compare && cap_a != kInf && zmin > imax(cap_a, cap_z + own);
			const bool reject = opaque_compare && cap_a != kInf && zlb >= 0 && zlb > imax(cap_a, cap_z + own);
			bool by_row = opaque_compare;
			for (int r = 0; r < 4 && by_row; r++)
				if (rmax[r] >= 0 && !(row_a[r] != kInf && rmin[r] > imax(row_a[r], row_z[r] + own)))
					by_row = false;
			dev->hiz[4] += ideal;
			dev->hiz[5] += zlb > zmin;
			bool passed = false;
			for (int l = 0; l < 32; l++)
			{
				if (!(cand[l] & DC_VALID))
					continue;
				dev->hiz_fragments[0]++;
				dev->hiz_fragments[2] += reject;
				if (defer_depth<SpecGeneric>(view[l], px[l], st, p.dz, p.dz_compressed, cand[l]))
				{
					passed = true;
					dev->hiz_fragments[1]++;
				}
			}
			dev->hiz[0]++;
			dev->hiz[1] += reject;
			dev->hiz[2] += !passed;
			dev->hiz[3] += reject && passed;
			dev->hiz[6] += by_row;
			dev->hiz[7] += by_row && passed;
			// what this entry may have written, for the later entries of the batch
			if (!by_row && (st.blend_flags & BF_Z_UPDATE))
			{
				const int wdz = 1 << int(p.dz_compressed);
				for (int r = 0; r < 4; r++)
				{
					if (rmax[r] < 0)
						continue;
					row_z[r] = imax(row_z[r], rmax[r]);
					if (wdz == 0x8000)
						row_a[r] = kInf;
					else if (row_a[r] != kInf)
						row_a[r] = imax(row_a[r], rmax[r] + 8 * imax(wdz << 1, 16));
				}
			}
			if (!reject && (st.blend_flags & BF_Z_UPDATE))
			{
				cap_z = imax(cap_z, zmax);
				const int written_dz = 1 << int(p.dz_compressed);
				if (written_dz == 0x8000)
					cap_a = kInf;
				else if (cap_a != kInf)
					cap_a = imax(cap_a, zmax + 8 * imax(written_dz << 1, 16));
			}
		}
	}
}

// How many (primitive, tile) entries the pass has under other tile shapes: per-entry work (list gather, mask phase,
// slot enumeration) scales with this count, lane utilisation with covered pixels / (entries * tile pixels).
static void hiz_tile_shape_stats(rdpb_device *dev, const PassBuffers &b, const FbInfo &fb)
{
	static const int shapes[4][2] = { { 8, 4 }, { 16, 4 }, { 8, 8 }, { 16, 8 } };
	for (uint32_t i = 0; i < fb.num_prims; i++)
	{
		const PrimSetup &p = b.prims[i];
		if (p.flags & SF_NO_PIXELS)
			continue;
		for (int s = 0; s < 4; s++)
		{
			const int tw = shapes[s][0], th = shapes[s][1];
			int row = -1, lo = 0, hi = -1;
			for (int y = imax(p.line_lo, 0); y <= imin(p.line_hi, int(fb.height) - 1); y++)
			{
				const Span &sp = b.spans[p.span_base + uint32_t(y - p.line_lo)];
				if (y / th != row)
				{
					if (hi >= lo)
						dev->tile_shapes[s] += uint64_t(hi / tw - lo / tw + 1);
					row = y / th;
					lo = 1 << 20;
					hi = -1;
				}
				if (!sp.valid)
					continue;
				const int xl = imax(int(sp.start_x), 0), xh = imin(int(sp.end_x), int(fb.width) - 1);
				if (xl > xh)
					continue;
				lo = imin(lo, xl);
				hi = imax(hi, xh);
				if (s == 0)
					dev->tile_shapes[4] += uint64_t(xh - xl + 1);
			}
			if (hi >= lo)
				dev->tile_shapes[s] += uint64_t(hi / tw - lo / tw + 1);
		}
	}
}

void rdpb_hostsim_fragment_stats(rdpb_device *dev, unsigned long long out[3])
{
	for (int i = 0; i < 3; i++)
		out[i] = dev->hiz_fragments[i];
}

void rdpb_hostsim_tile_shape_stats(rdpb_device *dev, unsigned long long out[5])
{
	for (int i = 0; i < 5; i++)
		out[i] = dev->tile_shapes[i];
}

void rdpb_hostsim_hiz_stats(rdpb_device *dev, unsigned long long out[8])
{
	for (int i = 0; i < 8; i++)
		out[i] = dev->hiz[i];
}

const char *rdpb_last_tile_variant(rdpb_device *dev) { return dev->last_deferred ? "generic+deferred" : "generic"; }

int rdpb_launch_tiles(rdpb_device *dev, const void *fb_info, const void *summary)
{
	if (g_null_compute)
		return 0;
	const FbInfo fb = *static_cast<const FbInfo *>(fb_info);
	if (summary && getenv("RDP_B200_DEBUG_PASSES"))
	{
		const PassSummary *sum = static_cast<const PassSummary *>(summary);
		fprintf(stderr, "[hostsim] pass: prims=%u fmt=%u states=%u rf_or=%08x rf_and=%08x bf_or=%02x bf_and=%02x tk=%04x uniform=%u "
		                "comb_rgb=%010llx comb_alpha=%06x blend_misc=%06x\n",
		        fb.num_prims, fb.fmt, sum->num_shaded_states, sum->rf_or, sum->rf_and, sum->bf_or, sum->bf_and, sum->texel_kinds,
		        sum->modes_uniform, (unsigned long long)sum->comb_rgb, sum->comb_alpha, sum->blend_misc);
	}
	const int sl = int(fb.scale_log2);
	PassBuffers b = sl ? dev->upscaled_buffers(fb) : dev->buffers();
	// RDP_HOSTSIM_DEFERRED=1: eligible passes take the scalar statement of the deferred tile pipeline instead of
	// the in-order walk (tests/test_hostsim_conformance.py runs the suites both ways).
	const bool deferred = !sl && summary && getenv("RDP_HOSTSIM_DEFERRED") && deferrable_pass(*static_cast<const PassSummary *>(summary), fb);
	dev->last_deferred = deferred;
	const bool hiz = !sl && summary && getenv("RDP_HOSTSIM_HIZ") && deferrable_pass(*static_cast<const PassSummary *>(summary), fb);
	if (hiz)
		hiz_tile_shape_stats(dev, b, fb);
	if (deferred)
		dev->deferred_passes++;
	for (uint32_t ty = 0; ty < fb.tiles_y; ty++)
	{
		for (uint32_t tx = 0; tx < fb.tiles_x; tx++)
		{
			const uint32_t tile = ty * fb.tiles_x + tx;
			bool any = false;
			for (uint32_t cw = 0; cw < fb.coarse_stride; cw++)
				any = any || b.coarse_mask[size_t(tile) * fb.coarse_stride + cw] != 0;
			if (!any)
				continue;
			if (hiz)
				hiz_check_tile(dev, b, fb, tile, tx, ty);
			for (int ly = 0; ly < int(kTileH); ly++)
			{
				for (int lx = 0; lx < int(kTileW); lx++)
				{
					const int x = int(tx) * int(kTileW) + lx, y = int(ty) * int(kTileH) + ly;
					if (x >= int(fb.width) || y >= int(fb.height))
						continue;
					FbView view;
					PixelState px;
					pass_pixel_begin<SpecGeneric>(b, fb, view, px, x, y, sl);
					dev->counters[1] += deferred ? pass_pixel_walk_deferred<SpecGeneric>(b, fb, view, px, x, y, tile)
					                             : pass_pixel_walk<SpecGeneric>(b, fb, view, px, x, y, tile);
					pass_pixel_end(fb, view, px, x, y, sl);
					uint32_t *word = nullptr;
					if (const uint32_t bits = pass_pixel_write_mask(b, fb, px, x, y, sl, word))
						*word |= bits;
				}
			}
			pass_tile_clear_masks(b, fb, tile);
			dev->counters[0]++;
		}
	}
	dev->launches++;
	return 0;
}

int rdpb_pass_end(rdpb_device *) { return 0; }

int rdpb_upscaling_init(rdpb_device *dev, uint32_t scale_log2)
{
	if (scale_log2 < 1 || scale_log2 > 3)
		return -1;
	dev->scale_log2 = scale_log2;
	const size_t slices = size_t(1) << (2 * scale_log2);
	dev->up_rdram.assign(slices * dev->rdram_size + size_t(kWriteMaskWords) * 4, 0); // + the super-sampled write mask
	dev->up_hidden.assign(slices * (dev->rdram_size >> 1), 3); // rdp_renderer.cpp:593
	dev->up_reference.assign(dev->rdram_size, 0);
	dev->up_spans.resize(size_t(kMaxSpansPerPass) << scale_log2);
	return 0;
}

int rdpb_launch_upscale_sync(rdpb_device *dev, const void *fb_info, int stage)
{
	if (g_null_compute || !dev->scale_log2)
		return 0;
	const FbInfo fb = *static_cast<const FbInfo *>(fb_info);
	UpscaleSync u;
	u.rdram = dev->rdram.data();
	u.hidden = dev->hidden.data();
	u.reference = dev->up_reference.data();
	u.up_rdram = dev->up_rdram.data();
	u.up_hidden = dev->up_hidden.data();
	u.rdram_size = dev->rdram_size;
	u.num_slices = 1u << (2 * dev->scale_log2);
	u.size_log2 = fb.fmt == FB_RGBA8888 ? 2u : ((fb.fmt == FB_RGBA5551 || fb.fmt == FB_IA88) ? 1u : 0u);
	u.fb_index = fb.color_addr >> u.size_log2;
	u.depth_index = fb.depth_addr >> 1;
	u.num_pixels = fb.width * fb.height;
	u.alias = fb.color_addr == fb.depth_addr;
	u.width = fb.width;
	u.dither = stage == RDPB_UPSCALE_SYNC_RESOLVE_DITHER;
	if (stage == RDPB_UPSCALE_SYNC_RESOLVE || stage == RDPB_UPSCALE_SYNC_RESOLVE_DITHER)
	{
		for (uint32_t y = 0; y < fb.height; y++)
			for (uint32_t x = 0; x < fb.width; x++)
				upscale_resolve_pixel(u, x, y);
		memset(dev->up_rdram.data() + size_t(u.num_slices) * dev->rdram_size, 0, size_t(kWriteMaskWords) * 4);
	}
	else
	{
		for (uint32_t i = 0; i < u.num_pixels; i++)
			upscale_sync_pixel(u, i, stage == RDPB_UPSCALE_SYNC_POST);
	}
	dev->launches++;
	return 0;
}

int rdpb_upscaled_read(rdpb_device *dev, void *rdram_dst, void *hidden_dst)
{
	if (!dev->scale_log2)
		return -1;
	if (rdram_dst)
		memcpy(rdram_dst, dev->up_rdram.data(), (size_t(1) << (2 * dev->scale_log2)) * dev->rdram_size);
	if (hidden_dst)
		memcpy(hidden_dst, dev->up_hidden.data(), dev->up_hidden.size());
	return 0;
}

int rdpb_launch_vi_scanout(rdpb_device *dev, const void *frame, const void *lines)
{
	if (g_null_compute)
		return 0;
	const ViFrame f = *static_cast<const ViFrame *>(frame);
	const size_t scale2 = size_t(1) << (2 * dev->scale_log2);
	const size_t max_out = size_t(640) * 576 * scale2;
	if (dev->vi_out.size() != max_out)
	{
		dev->vi_out.assign(max_out, 0);
		dev->vi_prev.assign(max_out, 0);
		dev->vi_stage[0].assign(size_t(kViStageWidth) * kViStageHeight * 2 * scale2, 0);
	}
	std::swap(dev->vi_out, dev->vi_prev);

	ViBuffers b;
	b.rdram = f.scale_log2 ? dev->up_rdram.data() : dev->rdram.data();
	b.hidden = f.scale_log2 ? dev->up_hidden.data() : dev->hidden.data();
	b.gamma_lut = dev->gamma_lut.data();
	b.lines = static_cast<const ViLine *>(lines);
	b.filtered = dev->vi_stage[0].data();
	b.out = dev->vi_out.data();
	b.prev = dev->vi_prev.data();

	for (int layer = 0; layer < (f.fetch_bug ? 2 : 1); layer++)
		for (int y = 0; y < f.filt_height; y++)
			for (int x = 0; x < f.filt_width; x++)
				vi_filter_thread(f, b, x, y, layer);
	for (int y = 0; y < f.out_height; y++)
		for (int x = 0; x < f.out_width; x++)
			vi_scale_thread(f, b, x, y);
	dev->launches += 2;

	dev->vi_result = dev->vi_out.data();
	if (f.scale_log2 && f.downscale_steps)
	{
		// Halve the finished image; the full-size one stays for the next field's merge.
		std::vector<uint32_t> *bufs[2] = { &dev->vi_stage[1], &dev->vi_stage[2] };
		const uint32_t *src = dev->vi_out.data();
		int w = f.out_width, h = f.out_height;
		for (uint32_t i = 0; i < f.downscale_steps; i++)
		{
			std::vector<uint32_t> &dst = *bufs[i & 1];
			dst.assign(size_t(w / 2) * (h / 2), 0);
			for (int y = 0; y < h / 2; y++)
				for (int x = 0; x < w / 2; x++)
					vi_downscale_thread(src, w, dst.data(), w / 2, x, y);
			src = dst.data();
			w /= 2;
			h /= 2;
			dev->launches++;
		}
		if (f.bob_scale)
		{
			std::vector<uint32_t> &dst = *bufs[f.downscale_steps & 1];
			const int out_h = f.bob_scale == 1u ? 2 * h : h;
			dst.assign(size_t(w) * out_h, 0);
			for (int y = 0; y < out_h; y++)
				for (int x = 0; x < w; x++)
					vi_bob_thread(src, w, h, dst.data(), f.bob_scale, f.bob_field != 0, x, y);
			src = dst.data();
			dev->launches++;
		}
		dev->vi_result = src;
	}
	return 0;
}

void *rdpb_vi_output_device_ptr(rdpb_device *dev) { return const_cast<uint32_t *>(dev->vi_result ? dev->vi_result : dev->vi_out.data()); }

int rdpb_vi_read_output(rdpb_device *dev, void *dst, uint32_t width, uint32_t height, int)
{
	memcpy(dst, dev->vi_result ? dev->vi_result : dev->vi_out.data(), size_t(width) * height * 4);
	return 0;
}

int rdpb_vi_read_output_fenced(rdpb_device *dev, void *dst, uint32_t width, uint32_t height, uint64_t)
{
	return rdpb_vi_read_output(dev, dst, width, height, 1);
}
int rdpb_vi_join(rdpb_device *) { return 0; }
int rdpb_signal(rdpb_device *, uint64_t) { return 0; }
int rdpb_wait(rdpb_device *, uint64_t) { return 0; }
int rdpb_poll(rdpb_device *, uint64_t) { return 1; }
void rdpb_submit_begin(rdpb_device *) {}
void rdpb_submit_end(rdpb_device *) {}
int rdpb_sync(rdpb_device *) { return 0; }
void *rdpb_stream(rdpb_device *) { return nullptr; }
uint64_t rdpb_launch_count(rdpb_device *dev) { return dev->launches; }

// Measurement plumbing: only the counters exist on the simulator; timing and replay are CUDA-only.
int rdpb_profile_enable(rdpb_device *, int, uint32_t) { return -1; }
int rdpb_profile_read(rdpb_device *, int, uint32_t *launches, double *total_ms)
{
	*launches = 0;
	*total_ms = 0.0;
	return -1;
}
int rdpb_counters_read(rdpb_device *dev, uint64_t *out8, int reset)
{
	for (int i = 0; i < 8; i++)
		out8[i] = dev->counters[i];
	if (reset)
		memset(dev->counters, 0, sizeof(dev->counters));
	return 0;
}
int rdpb_record_begin(rdpb_device *, size_t) { return -1; }
int rdpb_record_mark(rdpb_device *) { return -1; }
int rdpb_record_end(rdpb_device *) { return -1; }
uint32_t rdpb_record_steps(rdpb_device *) { return 0; }
int rdpb_replay_step(rdpb_device *, uint32_t) { return -1; }
int rdpb_replay_interleaved(rdpb_device *const *, uint32_t, uint32_t, uint32_t) { return -1; }
}
