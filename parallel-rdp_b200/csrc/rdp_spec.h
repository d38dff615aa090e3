// rdp_spec.h -- compile-time specialisation of the per-pixel pipeline over static render state.
//
// The reference compiles one Vulkan pipeline per static state through specialization constants, asynchronously,
// and runs an ubershader until the pipeline is ready (parallel-rdp/rdp_renderer.cpp:1721-1830, 3572-3579;
// shaders/rasterizer.comp:92-121).  Here the same split is done ahead of time with C++ templates: the tile kernel
// is instantiated for a table of hot static states (`Spec<...>` with exact mode words) and for broader feature
// classes (flag masks only), and a fully dynamic instantiation is the fallback.  A pass runs the first table
// entry whose constraints cover every non-fill state in it.
//
// How folding works: a Spec carries, for each flag word, a `may` mask (bits that can be set) and a `must` mask
// (bits that are set).  resolve_modes() computes (runtime & may) | must, so after inlining the compiler's
// known-bits analysis folds every `flags & BIT` test on a bit outside `may` or inside `must`; with `exact`
// the combiner / blender selectors are literal constants and their switch statements disappear.
#pragma once
#include "rdp_types.h"

namespace rdpb
{
// Texel fetch kinds, resolved per sample from (tile format, tile size, TLUT enable).
enum TexelKind
{
	TK_I4,     // 4-bit replicated (RGBA4 / I4)
	TK_IA4,
	TK_CI4,    // 4-bit index | palette, no TLUT
	TK_I8,     // 8-bit replicated (RGBA8 / CI8 / I8)
	TK_IA8,
	TK_RGBA16,
	TK_IA16,
	TK_RAW16,  // 16-bit word split into two bytes, replicated (the "32-bit CI/IA/I" oddities)
	TK_RGBA32,
	TK_YUV16,
	TK_CI4_TLUT,
	TK_CI8_TLUT,
	TK_CI16_TLUT,
	TK_COUNT
};

// (tile format, tile size, TLUT enable) -> texel kind; texture.h:498-600 only distinguishes these cases.
RDP_HD int resolve_texel_kind(int fmt, int size, bool tlut)
{
	if (tlut)
	{
		// texture.h:548-600: the TLUT path only distinguishes texel size.
		return size == 0 ? TK_CI4_TLUT : (size == 1 ? TK_CI8_TLUT : TK_CI16_TLUT);
	}
	switch (fmt)
	{
	case 0: // RGBA
		return size == 0 ? TK_I4 : (size == 1 ? TK_I8 : (size == 2 ? TK_RGBA16 : TK_RGBA32));
	case 1: // YUV
		return TK_YUV16;
	case 2: // CI
		return size == 0 ? TK_CI4 : (size == 1 ? TK_I8 : TK_RAW16);
	case 3: // IA
		return size == 0 ? TK_IA4 : (size == 1 ? TK_IA8 : (size == 2 ? TK_IA16 : TK_RAW16));
	default: // I (and the undefined formats 5..7 behave like I in the reference's switch fall-through: none sampled)
		return size == 0 ? TK_I4 : (size == 1 ? TK_I8 : TK_RAW16);
	}
}

// Static mode words of a state, packed (host: Renderer::finalize_state; device: Spec template arguments).
//   comb_rgb  : selector [cycle][i] at bits (cycle*4+i)*5, 5 bits each
//   comb_alpha: selector [cycle][i] at bits (cycle*4+i)*3, 3 bits each
//   blend_misc: blender [cycle][i] at bits (cycle*4+i)*2; dither_rgb 16-17, dither_alpha 18-19, cvg_mode 20-21, z_mode 22-23
RDP_HD uint64_t pack_comb_rgb(const uint8_t c[2][4])
{
	uint64_t v = 0;
	for (int i = 0; i < 8; i++)
		v |= uint64_t(c[i >> 2][i & 3] & 31u) << (i * 5);
	return v;
}
RDP_HD uint32_t pack_comb_alpha(const uint8_t c[2][4])
{
	uint32_t v = 0;
	for (int i = 0; i < 8; i++)
		v |= uint32_t(c[i >> 2][i & 3] & 7u) << (i * 3);
	return v;
}
RDP_HD uint32_t pack_blend_misc(const uint8_t b[2][4], uint32_t dither_rgb, uint32_t dither_alpha, uint32_t cvg_mode, uint32_t z_mode)
{
	uint32_t v = 0;
	for (int i = 0; i < 8; i++)
		v |= uint32_t(b[i >> 2][i & 3] & 3u) << (i * 2);
	return v | (dither_rgb << 16) | (dither_alpha << 18) | (cvg_mode << 20) | (z_mode << 22);
}

// Flags that never constrain a specialisation: they are consumed by set-up/span code or carry no per-pixel cost.
constexpr uint32_t kRfDontCare = RF_INTERLACE | RF_KEEP_ODD | RF_PRIM_DEPTH;

template <uint32_t RF_MAY, uint32_t RF_MUST, uint32_t BF_MAY, uint32_t BF_MUST, uint32_t FMT_MASK, uint32_t TK_MASK,
          bool EXACT = false, uint64_t COMB_RGB = 0, uint32_t COMB_ALPHA = 0, uint32_t BLEND_MISC = 0>
struct Spec
{
	static constexpr uint32_t rf_may = RF_MAY | RF_MUST | kRfDontCare;
	static constexpr uint32_t rf_must = RF_MUST;
	static constexpr uint32_t bf_may = BF_MAY | BF_MUST;
	static constexpr uint32_t bf_must = BF_MUST;
	static constexpr uint32_t fmt_mask = FMT_MASK; // bit f set: framebuffer format f may occur
	static constexpr uint32_t tk_mask = TK_MASK;   // bit k set: texel kind k may occur
	static constexpr bool exact = EXACT;
	static constexpr uint64_t comb_rgb = COMB_RGB;
	static constexpr uint32_t comb_alpha = COMB_ALPHA;
	static constexpr uint32_t blend_misc = BLEND_MISC;
	// Resident 128-thread blocks per SM the tile kernel is compiled for: 8 (<= 64 registers) when the state is exact,
	// 5 (<= 96 registers) for the feature classes and the uber-kernel.
	static constexpr int min_blocks_per_sm = EXACT ? 8 : 5;
	// The deferred tile kernel (k_tiles<S, true>) keeps ~5 KiB of fragment slots per warp in shared memory.
#ifndef RDPB_DEFERRED_BLOCKS_EXACT
#define RDPB_DEFERRED_BLOCKS_EXACT 8
#endif
	static constexpr int min_blocks_per_sm_deferred = EXACT ? RDPB_DEFERRED_BLOCKS_EXACT : 4;
	// Whether the blender can read shade alpha (blender input 1B = 2) or the dither pattern: what the deferred
	// kernel has to carry from its shade phase to its blend phase besides the combined colour.
	static constexpr bool blend_reads_aux = !EXACT || ((BF_MAY | BF_MUST) & BF_DITHER) != 0u || ((BLEND_MISC >> 2) & 3u) == 2u ||
	                                        ((BLEND_MISC >> 10) & 3u) == 2u;
};

// Fully dynamic: the uber-kernel fallback (ref shaders/ubershader.comp).
using SpecGeneric = Spec<~0u, 0u, ~0u, 0u, 0x1fu, ~0u>;

// Per-state modes as the per-pixel code consumes them.  Under an exact Spec every accessor is a literal; otherwise
// it reads the state snapshot on demand (nothing is materialised up front, which keeps the generic kernel's
// register count where it was).
template <class S>
struct StaticModes
{
	const PassState *st;
	uint32_t rf, bf;
	RDP_HD int comb_rgb(int cycle, int i) const
	{
		return S::exact ? int((S::comb_rgb >> ((cycle * 4 + i) * 5)) & 31u) : int(st->comb_rgb[cycle][i]);
	}
	RDP_HD int comb_alpha(int cycle, int i) const
	{
		return S::exact ? int((S::comb_alpha >> ((cycle * 4 + i) * 3)) & 7u) : int(st->comb_alpha[cycle][i]);
	}
	RDP_HD int blend(int cycle, int i) const
	{
		return S::exact ? int((S::blend_misc >> ((cycle * 4 + i) * 2)) & 3u) : int(st->blend_modes[cycle][i]);
	}
	RDP_HD int dither_rgb() const { return S::exact ? int((S::blend_misc >> 16) & 3u) : int(st->dither_rgb); }
	RDP_HD int dither_alpha() const { return S::exact ? int((S::blend_misc >> 18) & 3u) : int(st->dither_alpha); }
	RDP_HD int cvg_mode() const { return S::exact ? int((S::blend_misc >> 20) & 3u) : int(st->cvg_mode); }
	RDP_HD int z_mode() const { return S::exact ? int((S::blend_misc >> 22) & 3u) : int(st->z_mode); }
};

template <class S>
RDP_HD void resolve_modes(const PassState &st, StaticModes<S> &m)
{
	m.st = &st;
	m.rf = (st.raster_flags & S::rf_may) | S::rf_must;
	m.bf = (st.blend_flags & S::bf_may) | S::bf_must;
}

constexpr uint32_t ct_log2(uint32_t v) { return v <= 1u ? 0u : 1u + ct_log2(v >> 1); }
constexpr bool ct_single_bit(uint32_t v) { return v != 0u && (v & (v - 1u)) == 0u; }

template <class S>
RDP_HD uint32_t resolve_fb_fmt(uint32_t fmt)
{
	// A single possible format becomes a literal.
	return ct_single_bit(S::fmt_mask) ? ct_log2(S::fmt_mask) : fmt;
}

// What the host tells the backend about a pass so it can pick an instantiation (see backend_cuda.cu).
struct PassSummary
{
	uint32_t rf_or, rf_and;   // over the non-fill states that primitives of the pass use
	uint32_t bf_or, bf_and;
	uint32_t texel_kinds;     // bit k: texel kind k may be sampled
	uint32_t modes_uniform;   // all those states share one set of mode words (below)
	uint64_t comb_rgb;
	uint32_t comb_alpha, blend_misc;
	uint32_t num_shaded_states; // 0: the pass is fill-only
	uint32_t pad;
};
}
