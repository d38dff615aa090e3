// rdp_types.h -- plain-old-data shared by the C++ host, the CUDA kernels and the test-only host simulator.
//
// These are this project's own layouts (not the reference's Vulkan buffer layouts).  The host keeps
// per-primitive work tiny: it appends the *raw* RDP command words plus a 16-byte record that names a
// state snapshot; triangle/rectangle decode, derived constants and span set-up all run on the GPU.
//
// Behavioural reference for the fields: parallel-rdp/rdp_data_structures.hpp:95-254,
// parallel-rdp/rdp_common.hpp:37-322 (enums), parallel-rdp/rdp_device.cpp:268-860 (decode).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RDP_HD __host__ __device__ __forceinline__
#define RDP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define RDP_HD inline
#define RDP_HD_NOINLINE
#endif

namespace rdpb
{
// ---------------------------------------------------------------------------------------------
// Limits of one render pass (a pass = everything between two semantic flushes).
// The reference caps a pass at 256 primitives because of Vulkan-era buffer sizing
// (rdp_data_structures.hpp:364-377); only per-pixel primitive order is semantic, so we batch far more.
// ---------------------------------------------------------------------------------------------
enum : uint32_t
{
	kMaxPrimsPerPass = 16384,
	kMaxStatesPerPass = 1024,
	kMaxSpansPerPass = 1u << 20,
	kMaxUploadsPerPass = 255,
	kMaxPassWords = kMaxPrimsPerPass * 44u,
	kMaxFbWidth = 1024,
	kMaxFbHeight = 1024,
	kTileW = 8,
	kTileH = 4, // one warp = one 8x4 tile: 8 pixels x 4 B = one 32-byte sector per row of an RGBA32 target
	kMaxTilesX = kMaxFbWidth / kTileW,
	kMaxTilesY = kMaxFbHeight / kTileH,
	kTmemBytes = 4096,
};

// RDP opcodes (hardware command ids; reference: rdp_common.hpp:37-88).
enum Op : uint32_t
{
	OP_NOP = 0x00,
	// 0x01..0x07 are no-ops on hardware, two words like every other state command.  (The reference uses 1..4 as its
	// in-band Meta* commands; here meta operations are a ring record kind of their own, command_ring.hpp.)
	OP_TRI_FILL = 0x08,
	OP_TRI_FILL_Z = 0x09,
	OP_TRI_TEX = 0x0a,
	OP_TRI_TEX_Z = 0x0b,
	OP_TRI_SHADE = 0x0c,
	OP_TRI_SHADE_Z = 0x0d,
	OP_TRI_SHADE_TEX = 0x0e,
	OP_TRI_SHADE_TEX_Z = 0x0f,
	OP_TEX_RECT = 0x24,
	OP_TEX_RECT_FLIP = 0x25,
	OP_SYNC_LOAD = 0x26,
	OP_SYNC_PIPE = 0x27,
	OP_SYNC_TILE = 0x28,
	OP_SYNC_FULL = 0x29,
	OP_SET_KEY_GB = 0x2a,
	OP_SET_KEY_R = 0x2b,
	OP_SET_CONVERT = 0x2c,
	OP_SET_SCISSOR = 0x2d,
	OP_SET_PRIM_DEPTH = 0x2e,
	OP_SET_OTHER_MODES = 0x2f,
	OP_LOAD_TLUT = 0x30,
	OP_SET_TILE_SIZE = 0x32,
	OP_LOAD_BLOCK = 0x33,
	OP_LOAD_TILE = 0x34,
	OP_SET_TILE = 0x35,
	OP_FILL_RECT = 0x36,
	OP_SET_FILL_COLOR = 0x37,
	OP_SET_FOG_COLOR = 0x38,
	OP_SET_BLEND_COLOR = 0x39,
	OP_SET_PRIM_COLOR = 0x3a,
	OP_SET_ENV_COLOR = 0x3b,
	OP_SET_COMBINE = 0x3c,
	OP_SET_TEXTURE_IMAGE = 0x3d,
	OP_SET_MASK_IMAGE = 0x3e,
	OP_SET_COLOR_IMAGE = 0x3f,
};

// Framebuffer pixel formats (reference: shaders/fb_formats.h).
enum FbFmt : uint32_t
{
	FB_I4 = 0,
	FB_I8 = 1,
	FB_RGBA5551 = 2,
	FB_IA88 = 3,
	FB_RGBA8888 = 4,
};

// Raster ("shade") state flags.  Semantics follow StaticRasterizationFlagBits
// (rdp_data_structures.hpp:46-79); bit positions are our own.
enum RasterFlags : uint32_t
{
	RF_INTERLACE = 1u << 0,
	RF_KEEP_ODD = 1u << 1,
	RF_AA = 1u << 2,
	RF_PERSPECTIVE = 1u << 3,
	RF_TLUT = 1u << 4,
	RF_TLUT_IA = 1u << 5,
	RF_CVG_TIMES_ALPHA = 1u << 6,
	RF_ALPHA_CVG_SELECT = 1u << 7,
	RF_TWO_CYCLE = 1u << 8,
	RF_TEX_LOD = 1u << 9,
	RF_SHARPEN = 1u << 10,
	RF_DETAIL = 1u << 11,
	RF_FILL = 1u << 12,
	RF_COPY = 1u << 13,
	RF_SAMPLE_QUAD = 1u << 14,
	RF_ALPHA_TEST = 1u << 15,
	RF_ALPHA_TEST_DITHER = 1u << 16,
	RF_MID_TEXEL = 1u << 17,
	RF_USES_TEXEL0 = 1u << 18,
	RF_USES_TEXEL1 = 1u << 19,
	RF_USES_LOD = 1u << 20,
	RF_USES_PIPELINED_TEXEL1 = 1u << 21,
	RF_CONVERT_ONE = 1u << 22,
	RF_BILERP0 = 1u << 23,
	RF_BILERP1 = 1u << 24,
	RF_NOISE_DUAL = 1u << 25,
	RF_NOISE = 1u << 26,
	RF_PRIM_DEPTH = 1u << 27,
};

// Depth/blend state flags (semantics: DepthBlendFlagBits, rdp_data_structures.hpp:82-93).
enum BlendFlags : uint32_t
{
	BF_Z_COMPARE = 1u << 0,
	BF_Z_UPDATE = 1u << 1,
	BF_FORCE_BLEND = 1u << 2,
	BF_IMAGE_READ = 1u << 3,
	BF_COLOR_ON_CVG = 1u << 4,
	BF_TWO_CYCLE = 1u << 5,
	BF_AA = 1u << 6,
	BF_DITHER = 1u << 7,
};

// Triangle set-up flags (semantics: TriangleSetupFlagBits, rdp_data_structures.hpp:32-43).
enum SetupFlags : uint32_t
{
	SF_FLIP = 1u << 0,
	SF_DO_OFFSET = 1u << 1,
	SF_SKIP_XFRAC = 1u << 2,
	SF_INTERLACE = 1u << 3,
	SF_KEEP_ODD = 1u << 4,
	SF_FILL_COPY_RASTER = 1u << 5,
	SF_NO_PIXELS = 1u << 6, // never rasterises (Y clipped away); kept so the primitive counter advances
	// Internal upscaling renders this primitive at native resolution, replicated over each pixel's S*S sub-pixels: fill
	// rectangles and flipped texture rectangles always, texture rectangles under the native_resolution_tex_rect quirk
	// (ref rdp_device.cpp:730, 763, 806; shaders/shading.h:45-52).
	SF_DISABLE_UPSCALING = 1u << 7,
};

// Quirks a state snapshot carries (ref rdp_device.hpp:76-109).
enum QuirkFlags : uint32_t
{
	QUIRK_NATIVE_RESOLUTION_TEX_RECT = 1u << 0,
	QUIRK_NATIVE_TEXTURE_LOD = 1u << 1,
};

enum TileFlags : uint32_t
{
	TF_CLAMP_S = 1u << 0,
	TF_MIRROR_S = 1u << 1,
	TF_CLAMP_T = 1u << 2,
	TF_MIRROR_T = 1u << 3,
};

// One of the 8 RDP tile descriptors (SetTile + SetTileSize), 32 bytes.
struct alignas(16) TileDesc
{
	uint32_t slo, shi, tlo, thi; // 10.2 fixed point
	uint32_t offset, stride;     // TMEM bytes
	uint8_t fmt, size, palette, mask_s;
	uint8_t shift_s, mask_t, shift_t, flags;
};

// Everything non-geometric a primitive depends on, snapshotted by the host whenever any state
// command arrived since the previous primitive.  Primitives name a snapshot by index.
struct alignas(16) PassState
{
	uint32_t raster_flags;
	uint32_t blend_flags;
	uint8_t comb_rgb[2][4];   // [cycle] = {muladd(A), mulsub(B), mul(C), add(D)} selectors
	uint8_t comb_alpha[2][4];
	uint8_t blend_modes[2][4]; // [cycle] = {1a, 1b, 2a, 2b}
	uint8_t dither_rgb, dither_alpha, cvg_mode, z_mode;

	uint8_t prim_color[4], env_color[4], fog_color[4], blend_color[4];
	uint32_t fill_color;
	uint8_t key_center[3], min_lod;
	uint8_t key_scale[3], prim_lod_frac;
	int16_t convert[6]; // k0..k3 already as 2*sext9(k)+1, k4/k5 raw 9-bit (rdp_renderer.cpp:3512-3520)
	uint16_t prim_dz, pad0;
	int32_t prim_depth; // (z & 0x7fff) << 16

	uint16_t sc_xlo, sc_ylo, sc_xhi, sc_yhi; // scissor, 10.2
	uint32_t quirks; // QuirkFlags
	TileDesc tiles[8];
};

// Host -> device record, one per primitive, in submission order.
struct PrimRecord
{
	uint32_t word_offset; // first command word in the pass word buffer
	uint32_t span_base;   // first span slot (lines line_lo .. line_hi+1)
	uint16_t state_id;
	uint16_t tmem_instance; // TMEM snapshot index: number of uploads that preceded it in this pass
	int16_t line_lo, line_hi; // active scanlines after Y scissor; line_hi < line_lo => nothing to draw
};

// Decoded primitive, produced on the device by the set-up kernel.
struct alignas(16) PrimSetup
{
	int32_t xh, xm, xl;
	int32_t dxhdy, dxmdy, dxldy;
	int16_t yh, ym, yl;
	uint8_t flags, tile; // tile: bits 0-2 tile index, bits 3-5 max LOD level
	// rgba / stzw attribute planes: value, d/dx, d/de, d/dy
	int32_t rgba[4], drgba_dx[4], drgba_de[4], drgba_dy[4];
	int32_t stzw[4], dstzw_dx[4], dstzw_de[4], dstzw_dy[4];
	// Header: one aligned 16-byte word at byte 160 (read lane-parallel by the tile kernel's mask phase).
	uint32_t span_base;
	uint16_t state_id, tmem_instance;
	int16_t line_lo, line_hi;
	uint16_t dz;
	uint8_t dz_compressed, pad;
};
static_assert(offsetof(PrimSetup, span_base) == 160, "PrimSetup header must sit at byte 160");

// Per (primitive, scanline) edge-walk result, 64 bytes.
struct alignas(16) Span
{
	int32_t rgba[4];
	int32_t stzw[4];
	uint16_t xl[4]; // left edge per sub-scanline, x.3 fixed point, scissor-clamped
	uint16_t xr[4];
	// The next 8 bytes are what binning and the tile kernel's mask phase read: keep them in one aligned 8-byte word.
	int16_t start_x, end_x;
	int16_t lodlength;
	uint16_t valid;
	int32_t base_x; // integer X the attributes were latched at
	// Pixel range [lo, hi] (int16 | int16 << 16) this line can touch: what binning and the tile kernels' mask phase use.
	// Equal to start_x / end_x except for SF_DISABLE_UPSCALING primitives in an upscaled pass, where it is the range of
	// upscaled pixels whose native pixel's first sub-pixel lies in [start_x, end_x].
	uint32_t touch_range;
};
static_assert(offsetof(Span, start_x) == 48, "Span header word must sit at byte 48");

// TMEM upload descriptor (LoadTile / LoadBlock / LoadTLUT), decomposed on the host.
// Semantics: UploadInfo, rdp_data_structures.hpp:182-202.
struct TmemUpload
{
	int32_t width, height;
	float min_t_mod, max_t_mod;
	int32_t vram_addr, vram_width, vram_size, vram_effective_width;
	int32_t tmem_offset, tmem_stride_words, tmem_size, tmem_fmt;
	int32_t mode; // 0 tile, 1 tlut, 2 block
	float inv_tmem_stride_words;
	int32_t dxt;
	int32_t pad;
};

// Framebuffer binding for a pass.
struct FbInfo
{
	uint32_t color_addr; // byte address in RDRAM
	uint32_t depth_addr;
	uint32_t width;      // pixels
	uint32_t height;     // deduced from the primitives of the pass
	uint32_t fmt;        // FbFmt
	uint32_t rdram_size; // bytes, power of two
	uint32_t base_prim_index; // global primitive counter at pass start (noise seed)
	uint32_t num_prims;
	uint32_t mask_stride;   // 32-bit words of fine mask per tile = ceil(num_prims/32)
	uint32_t coarse_stride; // 32-bit words of coarse mask per tile = ceil(mask_stride/32)
	uint32_t tiles_x, tiles_y;
	uint32_t num_jobs; // span jobs = sum over primitives of span_job_count()
	// log2 of the internal upscaling factor S of this pass.  When non-zero, width / height / tiles_x / tiles_y are in
	// upscaled pixels, the pass reads and writes the upscaled RDRAM domain (S*S slices) and its own span / mask buffers;
	// addresses, primitive records and scanline ranges stay native (ref rdp_renderer.cpp:1865-1868, 2138-2141).
	uint32_t scale_log2;
	// Super-sampled read-back: the upscaled pass also records which native pixels it stored to (2 bits per native
	// pixel, 4x4 pixels per word, behind the last RDRAM slice; ref memory_interfacing.h:305-337).
	uint32_t write_mask;
	// ceil(2^32 / tiles_x), 0 when tiles_x == 1: tile / tiles_x == umulhi(tile, tile_row_recip) for every tile index below
	// 2^32 / tiles_x.  Filled in by the CUDA backend at launch (the tile kernels read it from the constant bank, so the
	// tile index -> (column, row) split costs a multiply instead of a division per tile); zero elsewhere.
	uint32_t tile_row_recip;
};

// Which passes are worth rendering upscaled: the reference's heuristic (rdp_renderer.cpp:3000-3012: only 16/32-bit
// colour images at least 256 pixels wide can be what reaches the screen), plus our own range limit -- a Span keeps
// its X edges as 16-bit x.3 values, so the upscaled width has to stay below 8192 pixels.
RDP_HD bool pass_wants_upscaling(const FbInfo &fb, uint32_t scale_log2)
{
	return (fb.fmt == FB_RGBA5551 || fb.fmt == FB_RGBA8888) && fb.width >= 256u && (fb.width << scale_log2) < 8192u;
}
// With super-sampled read-back every pass is rendered upscaled only (rdp_renderer.cpp:2955, 2990-2991); what remains
// is our own range limit.
RDP_HD bool pass_can_super_sample(const FbInfo &fb, uint32_t scale_log2)
{
	return (fb.width << scale_log2) < 8192u;
}
// Words of the super-sampled write mask (2 bits per native pixel of the largest framebuffer, rdp_renderer.cpp:581-582).
enum : uint32_t { kWriteMaskWords = 2u * 1024u * 1024u / 32u };

// A span job is one (primitive, 32-scanline chunk): one warp of the span/bin kernel.  Chunks start at the
// primitive's first tile row so that every 4-lane group of a warp owns exactly one tile row.
enum : uint32_t { kSpanJobLines = 32, kMaxJobsPerPass = (kMaxSpansPerPass / kSpanJobLines) + kMaxPrimsPerPass };
RDP_HD uint32_t span_job_count(int line_lo, int line_hi)
{
	if (line_hi < line_lo)
		return 0;
	const int first = line_lo & ~int(kTileH - 1);
	return uint32_t((line_hi + 2 - first + int(kSpanJobLines) - 1) / int(kSpanJobLines));
}
RDP_HD uint32_t span_job_pack(uint32_t prim, uint32_t chunk) { return prim | (chunk << 16); }

// ---------------------------------------------------------------------------------------------
// VI
// ---------------------------------------------------------------------------------------------
struct ViLine
{
	int32_t h_start, h_start_clamp, h_end_clamp;
	int32_t x_start, x_add;
	int32_t y_start, y_add, y_base;
};

struct ViRegs
{
	int32_t vi_width, vi_offset, v_current_line;
	int32_t is_pal;
	uint32_t status;
	int32_t init_y_add;
	int32_t h_start_clamp, h_res_clamp;
	int32_t h_start, h_res;
	int32_t v_start, v_res;
	int32_t max_x, max_y;
};

static_assert(sizeof(TileDesc) == 32, "TileDesc must be 32 bytes");
static_assert(sizeof(PrimRecord) == 16, "PrimRecord must be 16 bytes");
static_assert(sizeof(Span) == 64, "Span must be 64 bytes");
static_assert(sizeof(TmemUpload) == 64, "TmemUpload must be 64 bytes");
static_assert((sizeof(PrimSetup) & 15) == 0, "PrimSetup must be 16-byte aligned");
static_assert((sizeof(PassState) & 15) == 0, "PassState must be 16-byte aligned");
}
