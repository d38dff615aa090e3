/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Thin C driver around the reference's own CPU renderer (Angrylion-Plus as bundled by
 * paraLLEl-RDP).  The renderer sources are compiled where they lie under
 * /root/reference/angrylion-rdp-plus/src/core (see oracle/Makefile); this file only supplies
 * the host callbacks the core expects (vdac_*, msg_*) and a flat C API for ctypes.
 *
 * It drives the core the same way the reference's harness does
 * (reference: replayer_driver_angrylion.cpp:217-252): parallel=false, VI_MODE_NORMAL,
 * VI_INTERP_LINEAR, DP_COMPAT_HIGH, one rdp_cmd(0, words) per command and
 * n64video_update_screen() per frame.
 */
#include "n64video.h"
#include "vdac.h"
#include "msg.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void rdp_cmd(uint32_t wid, const uint32_t *args);
extern uint8_t rdram_hidden[RDRAM_MAX_SIZE / 2];
extern uint8_t *get_tmem(void);

static struct n64video_config cfg;
static uint8_t *g_rdram;
static uint32_t g_rdram_size;
static uint32_t g_vi_regs[VI_NUM_REG];
static uint32_t g_dp_regs[DP_NUM_REG];
static uint32_t *g_vi_ptrs[VI_NUM_REG];
static uint32_t *g_dp_ptrs[DP_NUM_REG];
static uint32_t g_irq;
static int g_verbose;

static uint32_t *g_frame;
static uint32_t g_frame_cap;
static uint32_t g_frame_w, g_frame_h;
static int g_frame_valid;

static void irq_cb(void) {}

/* ---- callbacks required by the core (harness equivalents: replayer_driver_angrylion.cpp:120-167) ---- */
void vdac_init(struct n64video_config *c) { (void)c; }
void vdac_close(void) {}
void vdac_read(struct frame_buffer *fb, bool alpha) { (void)fb; (void)alpha; }

void vdac_write(struct frame_buffer *fb)
{
    uint32_t n = fb->width * fb->height;
    if (n > g_frame_cap) {
        free(g_frame);
        g_frame = (uint32_t *)malloc((size_t)n * 4);
        g_frame_cap = n;
    }
    for (uint32_t y = 0; y < fb->height; y++)
        memcpy(g_frame + (size_t)y * fb->width, fb->pixels + (size_t)y * fb->pitch, (size_t)fb->width * 4);
    g_frame_w = fb->width;
    g_frame_h = fb->height;
    g_frame_valid = 1;
}

void vdac_sync(bool invalid)
{
    if (invalid) {
        g_frame_w = g_frame_h = 0;
        g_frame_valid = 0;
    }
}

static void vmsg(const char *tag, const char *fmt, va_list va)
{
    if (!g_verbose)
        return;
    fprintf(stderr, "[angrylion %s] ", tag);
    vfprintf(stderr, fmt, va);
    fputc('\n', stderr);
}
void msg_error(const char *e, ...)   { va_list va; va_start(va, e); g_verbose = 1; vmsg("error", e, va); va_end(va); }
void msg_warning(const char *e, ...) { va_list va; va_start(va, e); vmsg("warn", e, va); va_end(va); }
void msg_debug(const char *e, ...)   { va_list va; va_start(va, e); vmsg("debug", e, va); va_end(va); }

/* ---- flat API ---- */
int ref_init(uint32_t rdram_size, int parallel, uint32_t num_workers)
{
    if (rdram_size > RDRAM_MAX_SIZE)
        return -1;
    if (!g_rdram || g_rdram_size != rdram_size) {
        free(g_rdram);
        g_rdram = (uint8_t *)calloc(1, rdram_size);
        g_rdram_size = rdram_size;
    } else
        memset(g_rdram, 0, rdram_size);

    for (unsigned i = 0; i < VI_NUM_REG; i++) g_vi_ptrs[i] = &g_vi_regs[i];
    for (unsigned i = 0; i < DP_NUM_REG; i++) g_dp_ptrs[i] = &g_dp_regs[i];
    memset(g_vi_regs, 0, sizeof(g_vi_regs));
    memset(g_dp_regs, 0, sizeof(g_dp_regs));

    memset(&cfg, 0, sizeof(cfg));
    cfg.gfx.rdram = g_rdram;
    cfg.gfx.rdram_size = rdram_size;
    cfg.gfx.vi_reg = g_vi_ptrs;
    cfg.gfx.dp_reg = g_dp_ptrs;
    cfg.gfx.mi_intr_reg = &g_irq;
    cfg.gfx.mi_intr_cb = irq_cb;
    cfg.vi.mode = VI_MODE_NORMAL;
    cfg.vi.interp = VI_INTERP_LINEAR;
    cfg.dp.compat = DP_COMPAT_HIGH;
    cfg.parallel = parallel != 0;
    cfg.num_workers = num_workers;
    n64video_init(&cfg);
    memset(get_tmem(), 0, 0x1000);
    return 0;
}

void ref_close(void) { n64video_close(); }
void ref_set_verbose(int v) { g_verbose = v; }
uint8_t *ref_rdram(void) { return g_rdram; }
uint32_t ref_rdram_size(void) { return g_rdram_size; }
uint8_t *ref_hidden_rdram(void) { return rdram_hidden; }
uint32_t ref_hidden_rdram_size(void) { return (uint32_t)sizeof(rdram_hidden); }
uint8_t *ref_tmem(void) { return get_tmem(); }

/* Command length in 32-bit words, by opcode (reference: angrylion rdp.c:391-462 length table). */
static unsigned cmd_words(uint32_t w0)
{
    switch ((w0 >> 24) & 63) {
    case 0x08: return 8;   case 0x09: return 12;
    case 0x0a: return 24;  case 0x0b: return 28;
    case 0x0c: return 24;  case 0x0d: return 28;
    case 0x0e: return 40;  case 0x0f: return 44;
    case 0x24: case 0x25: return 4;
    default: return 2;
    }
}

void ref_command(const uint32_t *words) { rdp_cmd(0, words); }

/* Run a packed stream of commands; returns number of commands executed. */
uint32_t ref_command_stream(const uint32_t *words, uint32_t num_words)
{
    uint32_t pos = 0, n = 0;
    uint32_t tmp[44];
    while (pos < num_words) {
        unsigned len = cmd_words(words[pos]);
        if (pos + len > num_words)
            break;
        memcpy(tmp, words + pos, len * 4);
        rdp_cmd(0, tmp);
        pos += len;
        n++;
    }
    return n;
}

/* Feed a packed command stream the way real hardware does: as a display list in RDRAM consumed by
 * n64video_process_list() (reference: angrylion n64video.c:230-310).  This is the only route that engages the
 * oracle's scanline-parallel worker pool (config.parallel), so the multi-core CPU baseline uses it.
 * `list_addr` must be 8-byte aligned and the list must fit below rdram_size. */
int ref_process_list(const uint32_t *words, uint32_t num_words, uint32_t list_addr)
{
    if ((list_addr & 7u) || (uint64_t)list_addr + (uint64_t)num_words * 4u > g_rdram_size)
        return -1;
    memcpy(g_rdram + list_addr, words, (size_t)num_words * 4u); /* RDRAM is stored as host-endian 32-bit words */
    g_dp_regs[DP_STATUS] = 0; /* no XBUS DMA: read the list from RDRAM */
    g_dp_regs[DP_START] = list_addr;
    g_dp_regs[DP_CURRENT] = list_addr;
    g_dp_regs[DP_END] = list_addr + num_words * 4u;
    n64video_process_list();
    return 0;
}

void ref_set_vi_register(uint32_t reg, uint32_t value)
{
    if (reg < VI_NUM_REG)
        g_vi_regs[reg] = value;
}

/* Runs the VI; returns 1 if a frame was produced. */
int ref_update_screen(void)
{
    g_frame_valid = 0;
    g_frame_w = g_frame_h = 0;
    n64video_update_screen();
    return g_frame_valid;
}

const uint32_t *ref_frame(uint32_t *w, uint32_t *h)
{
    *w = g_frame_w;
    *h = g_frame_h;
    return g_frame_valid ? g_frame : NULL;
}
