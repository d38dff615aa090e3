// rdp_vi.cuh -- VI scan-out stages (placeholder until the RDP path is parity-green).
#pragma once
#include "rdp_types.h"

namespace rdpb
{
enum : uint32_t
{
	kViStageWidth = 1024,
	kViStageHeight = 640,
};

struct ViContext
{
	const uint8_t *rdram;
	const uint8_t *hidden;
	uint32_t rdram_size;
	const uint8_t *gamma_lut;
	ViRegs regs;
	const ViLine *lines;
	uint32_t flags;
	uint32_t frame_count;
	uint32_t out_width, out_height;
	uint32_t *stage0, *stage1, *stage2;
	uint32_t *out;
	const uint32_t *prev;
};

void build_blender_lut(uint8_t *lut);
void build_gamma_lut(uint8_t *lut);

inline void vi_run_all_stages_serial(const ViContext &) {}
}
