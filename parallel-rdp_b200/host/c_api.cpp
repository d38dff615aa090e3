// c_api.cpp -- extern "C" wrappers declared in include/rdp_b200.h over RDP::CommandProcessor.
#include "rdp_b200.h"
#include "rdp_device.hpp"
#include "rdp_dump.hpp"
#include <string.h>

using namespace RDP;

struct rdp_b200_processor
{
	CommandProcessor impl;
	// The C callback behind the reference's ValidationInterface object.
	struct Validation final : ValidationInterface
	{
		rdp_b200_validation_callback callback = nullptr;
		void *user = nullptr;
		void report_rdp_crash(ValidationError err, const char *msg) override
		{
			if (callback)
				callback(user, int(err), msg);
		}
	} validation;
	rdp_b200_processor(int dev, void *rdram, size_t off, size_t size, size_t hidden, uint32_t flags)
		: impl(DeviceConfig{dev}, rdram, off, size, hidden, flags)
	{
	}
};

struct rdp_b200_dump_player
{
	DumpPlayer impl;
};
struct rdp_b200_dump_writer
{
	RDPDumpWriter impl;
};

static ScanoutOptions convert_options(const rdp_b200_scanout_options *o)
{
	ScanoutOptions opts;
	if (!o)
		return opts;
	opts.crop_overscan_pixels = o->crop_overscan_pixels;
	opts.crop_rect.left = o->crop_left;
	opts.crop_rect.right = o->crop_right;
	opts.crop_rect.top = o->crop_top;
	opts.crop_rect.bottom = o->crop_bottom;
	opts.crop_rect.enable = o->crop_enable != 0;
	opts.downscale_steps = o->downscale_steps;
	opts.persist_frame_on_invalid_input = o->persist_frame_on_invalid_input != 0;
	opts.blend_previous_frame = o->blend_previous_frame != 0;
	opts.upscale_deinterlacing = o->upscale_deinterlacing != 0;
	opts.vi.aa = o->vi_aa != 0;
	opts.vi.scale = o->vi_scale != 0;
	opts.vi.serrate = o->vi_serrate != 0;
	opts.vi.dither_filter = o->vi_dither_filter != 0;
	opts.vi.divot_filter = o->vi_divot_filter != 0;
	opts.vi.gamma_dither = o->vi_gamma_dither != 0;
	return opts;
}

extern "C" {

rdp_b200_processor *rdp_b200_create(int cuda_device, void *rdram_ptr, size_t rdram_offset, size_t rdram_size,
                                    size_t hidden_rdram_size, uint32_t flags)
{
	return new rdp_b200_processor(cuda_device, rdram_ptr, rdram_offset, rdram_size, hidden_rdram_size, flags);
}

void rdp_b200_destroy(rdp_b200_processor *p) { delete p; }
int rdp_b200_device_is_supported(const rdp_b200_processor *p) { return p && p->impl.device_is_supported() ? 1 : 0; }

void rdp_b200_enqueue_command(rdp_b200_processor *p, unsigned n, const uint32_t *w) { p->impl.enqueue_command(n, w); }
void rdp_b200_enqueue_command_direct(rdp_b200_processor *p, unsigned n, const uint32_t *w) { p->impl.enqueue_command_direct(n, w); }
void rdp_b200_enqueue_command_stream(rdp_b200_processor *p, size_t n, const uint32_t *w) { p->impl.enqueue_command_stream(n, w); }
void rdp_b200_enqueue_command_stream_borrowed(rdp_b200_processor *p, size_t n, const uint32_t *w) { p->impl.enqueue_command_stream_borrowed(n, w); }

void rdp_b200_flush(rdp_b200_processor *p) { p->impl.flush(); }
uint64_t rdp_b200_signal_timeline(rdp_b200_processor *p) { return p->impl.signal_timeline(); }
void rdp_b200_wait_for_timeline(rdp_b200_processor *p, uint64_t index) { p->impl.wait_for_timeline(index); }
void rdp_b200_idle(rdp_b200_processor *p) { p->impl.idle(); }
void rdp_b200_begin_frame_context(rdp_b200_processor *p) { p->impl.begin_frame_context(); }

void rdp_b200_set_validation_callback(rdp_b200_processor *p, rdp_b200_validation_callback callback, void *user)
{
	p->impl.idle(); // the command thread reads the interface pointer
	p->validation.callback = callback;
	p->validation.user = user;
	p->impl.set_validation_interface(callback ? &p->validation : nullptr);
}

void rdp_b200_set_quirks(rdp_b200_processor *p, int native_resolution_tex_rect, int native_texture_lod)
{
	Quirks q;
	q.set_native_resolution_tex_rect(native_resolution_tex_rect != 0);
	q.set_native_texture_lod(native_texture_lod != 0);
	p->impl.set_quirks(q);
}

void *rdp_b200_begin_read_rdram(rdp_b200_processor *p) { return p->impl.begin_read_rdram(); }
void rdp_b200_end_write_rdram(rdp_b200_processor *p) { p->impl.end_write_rdram(); }
void rdp_b200_mark_rdram_written(rdp_b200_processor *p, size_t offset, size_t size) { p->impl.mark_rdram_written(offset, size); }
void rdp_b200_rdram_mirror_stats(rdp_b200_processor *p, uint64_t out4[4]) { p->impl.get_rdram_mirror_stats(out4); }
void *rdp_b200_begin_read_hidden_rdram(rdp_b200_processor *p) { return p->impl.begin_read_hidden_rdram(); }
void rdp_b200_end_write_hidden_rdram(rdp_b200_processor *p) { p->impl.end_write_hidden_rdram(); }
size_t rdp_b200_get_rdram_size(const rdp_b200_processor *p) { return p->impl.get_rdram_size(); }
size_t rdp_b200_get_hidden_rdram_size(const rdp_b200_processor *p) { return p->impl.get_hidden_rdram_size(); }
void *rdp_b200_get_tmem(rdp_b200_processor *p) { return p->impl.get_tmem(); }

void rdp_b200_set_vi_register(rdp_b200_processor *p, uint32_t reg, uint32_t value)
{
	if (reg < uint32_t(VIRegister::Count))
		p->impl.set_vi_register(VIRegister(reg), value);
}
void rdp_b200_set_vi_registers(rdp_b200_processor *p, const uint32_t values[14])
{
	for (uint32_t r = 0; r < 14; r++)
		p->impl.set_vi_register(VIRegister(r), values[r]);
}
void rdp_b200_begin_vi_register_per_scanline(rdp_b200_processor *p, uint32_t flags) { p->impl.begin_vi_register_per_scanline(flags); }
void rdp_b200_set_vi_register_for_scanline(rdp_b200_processor *p, uint32_t reg_bit, uint32_t value)
{
	p->impl.set_vi_register_for_scanline(VideoInterface::PerScanlineRegisterBits(reg_bit), value);
}
void rdp_b200_latch_vi_register_for_scanline(rdp_b200_processor *p, unsigned vi_line) { p->impl.latch_vi_register_for_scanline(vi_line); }
void rdp_b200_end_vi_register_per_scanline(rdp_b200_processor *p) { p->impl.end_vi_register_per_scanline(); }

void rdp_b200_default_scanout_options(rdp_b200_scanout_options *o)
{
	memset(o, 0, sizeof(*o));
	o->upscale_deinterlacing = 1;
	o->vi_aa = o->vi_scale = o->vi_serrate = o->vi_dither_filter = o->vi_divot_filter = o->vi_gamma_dither = 1;
}

const void *rdp_b200_scanout(rdp_b200_processor *p, const rdp_b200_scanout_options *opts, unsigned *width, unsigned *height)
{
	ScanoutImage img = p->impl.scanout(convert_options(opts));
	*width = img.width;
	*height = img.height;
	// The contract of this entry point: the pointer is good in rdp_b200_cuda_stream() order.
	if (p->impl.get_backend_device())
		rdpb_vi_join(p->impl.get_backend_device());
	return img.device_rgba8;
}

void rdp_b200_scanout_sync(rdp_b200_processor *p, const rdp_b200_scanout_options *opts, uint32_t *rgba,
                           size_t capacity_pixels, unsigned *width, unsigned *height)
{
	VIScanoutBuffer buf;
	p->impl.scanout_async_buffer(buf, convert_options(opts));
	*width = buf.width;
	*height = buf.height;
	size_t n = size_t(buf.width) * buf.height;
	if (n > capacity_pixels)
		n = capacity_pixels;
	if (n)
		memcpy(rgba, buf.pixels.data(), n * 4);
}

uint64_t rdp_b200_scanout_async(rdp_b200_processor *p, const rdp_b200_scanout_options *opts, uint32_t *rgba, size_t capacity_pixels,
                                unsigned *width, unsigned *height)
{
	return p->impl.scanout_async(rgba, capacity_pixels, width, height, convert_options(opts));
}
void rdp_b200_scanout_wait(rdp_b200_processor *p, uint64_t fence) { p->impl.scanout_wait(fence); }

const char *rdp_b200_backend_name(void) { return rdpb_backend_name(); }
uint64_t rdp_b200_launch_count(rdp_b200_processor *p) { return p->impl.get_backend_device() ? rdpb_launch_count(p->impl.get_backend_device()) : 0; }
uint64_t rdp_b200_primitive_count(rdp_b200_processor *p) { return p->impl.get_primitive_counter(); }
uint64_t rdp_b200_pass_count(rdp_b200_processor *p) { return p->impl.get_pass_counter(); }
void rdp_b200_host_timing(rdp_b200_processor *p, double out4[4]) { p->impl.get_host_timing(out4); }
void *rdp_b200_backend_device(rdp_b200_processor *p) { return p->impl.get_backend_device(); }
int rdp_b200_set_command_thread_affinity(rdp_b200_processor *p, const int *cpus, unsigned num_cpus)
{
#if defined(__linux__)
	p->impl.set_command_thread_affinity(std::vector<int>(cpus, cpus + num_cpus));
	return 0;
#else
	(void)p; (void)cpus; (void)num_cpus;
	return -1;
#endif
}
const char *rdp_b200_last_tile_variant(rdp_b200_processor *p) { return p->impl.get_backend_device() ? rdpb_last_tile_variant(p->impl.get_backend_device()) : ""; }
void *rdp_b200_cuda_stream(rdp_b200_processor *p) { return p->impl.get_backend_device() ? rdpb_stream(p->impl.get_backend_device()) : nullptr; }

// ---- internal upscaling ----
unsigned rdp_b200_upscaling_factor(rdp_b200_processor *p) { return 1u << p->impl.get_upscaling_log2(); }
uint64_t rdp_b200_upscaled_pass_count(rdp_b200_processor *p) { return p->impl.get_upscaled_pass_count(); }
int rdp_b200_read_upscaled_rdram(rdp_b200_processor *p, void *rdram_dst, void *hidden_dst)
{
	p->impl.idle();
	return p->impl.get_backend_device() ? rdpb_upscaled_read(p->impl.get_backend_device(), rdram_dst, hidden_dst) : -1;
}

// ---- RDPDUMP2 ------------------------------------------------------------------------------------------------
rdp_b200_dump_player *rdp_b200_dump_open(const char *path)
{
	auto *d = new rdp_b200_dump_player;
	if (!d->impl.load_dump(path))
	{
		delete d;
		return nullptr;
	}
	return d;
}
void rdp_b200_dump_close(rdp_b200_dump_player *d) { delete d; }
size_t rdp_b200_dump_rdram_size(const rdp_b200_dump_player *d) { return d->impl.get_rdram_size(); }
size_t rdp_b200_dump_hidden_rdram_size(const rdp_b200_dump_player *d) { return d->impl.get_hidden_rdram_size(); }
int rdp_b200_dump_rewind(rdp_b200_dump_player *d) { return d->impl.rewind() ? 1 : 0; }

int rdp_b200_dump_next(rdp_b200_dump_player *d, rdp_b200_dump_event *ev)
{
	DumpEvent e;
	if (!d->impl.next_event(e))
		return 0;
	ev->kind = uint32_t(e.kind);
	ev->arg0 = e.arg0;
	ev->arg1 = e.arg1;
	ev->num_words = e.num_words;
	ev->words = e.words;
	ev->data = e.data;
	ev->size = e.size;
	return 1;
}

long rdp_b200_dump_replay(rdp_b200_dump_player *d, rdp_b200_processor *p, rdp_b200_dump_event *stop)
{
	long commands = 0;
	rdp_b200_dump_event ev;
	while (rdp_b200_dump_next(d, &ev))
	{
		switch (ev.kind)
		{
		case RDP_B200_DUMP_RDP_COMMAND:
			p->impl.enqueue_command(ev.num_words, ev.words);
			commands++;
			break;
		case RDP_B200_DUMP_SET_VI_REGISTER:
			p->impl.set_vi_register(VIRegister(ev.arg0), ev.arg1);
			break;
		case RDP_B200_DUMP_UPDATE_RDRAM:
			p->impl.idle();
			memcpy(p->impl.begin_read_rdram(), ev.data, ev.size < p->impl.get_rdram_size() ? ev.size : p->impl.get_rdram_size());
			p->impl.end_write_rdram();
			break;
		case RDP_B200_DUMP_UPDATE_HIDDEN_RDRAM:
			p->impl.idle();
			memcpy(p->impl.begin_read_hidden_rdram(), ev.data,
			       ev.size < p->impl.get_hidden_rdram_size() ? ev.size : p->impl.get_hidden_rdram_size());
			p->impl.end_write_hidden_rdram();
			break;
		case RDP_B200_DUMP_SIGNAL_COMPLETE:
			p->impl.flush();
			/* fall through */
		default: // END_FRAME, EOF
			if (stop)
				*stop = ev;
			return commands;
		}
	}
	return -1;
}

rdp_b200_dump_writer *rdp_b200_dump_writer_open(const char *path, uint32_t rdram_size, uint32_t hidden_rdram_size)
{
	auto *w = new rdp_b200_dump_writer;
	if (!w->impl.init(path, rdram_size, hidden_rdram_size))
	{
		delete w;
		return nullptr;
	}
	return w;
}
void rdp_b200_dump_writer_flush_dram(rdp_b200_dump_writer *w, const void *rdram, uint32_t size) { w->impl.flush_dram(rdram, size); }
void rdp_b200_dump_writer_flush_hidden_dram(rdp_b200_dump_writer *w, const void *hidden, uint32_t size) { w->impl.flush_hidden_dram(hidden, size); }
void rdp_b200_dump_writer_command(rdp_b200_dump_writer *w, uint32_t op, const uint32_t *words, uint32_t n) { w->impl.emit_command(op, words, n); }
void rdp_b200_dump_writer_set_vi_register(rdp_b200_dump_writer *w, uint32_t reg, uint32_t value) { w->impl.set_vi_register(reg, value); }
void rdp_b200_dump_writer_end_frame(rdp_b200_dump_writer *w) { w->impl.end_frame(); }
void rdp_b200_dump_writer_signal_complete(rdp_b200_dump_writer *w) { w->impl.signal_complete(); }
void rdp_b200_dump_writer_close(rdp_b200_dump_writer *w) { delete w; }
}
