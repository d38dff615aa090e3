// farm_driver.cpp -- native driver loop for bench.py's end-to-end legs and the stream farm (BASELINE C5).
//
// Measurement plumbing, not product: it calls nothing but the public C-ABI of include/rdp_b200.h, the way an emulator
// farm would -- a few host threads, each feeding several independent processors: borrowed command words from pinned
// host memory in, VI registers, an asynchronous scan-out with the image copied back to pinned host memory, a bounded
// number of frames in flight per processor.  Python starts one farm_run() per timed region (ctypes drops the GIL),
// so no interpreter time sits between two frames.
#include "rdp_b200.h"

#include <atomic>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>

extern "C" {

enum
{
	FARM_READ_BACK_RDRAM = 1u << 0, // every frame also signals the timeline and waits for it before the frame slot is reused:
	                                // the rendered colour + depth images travel back into the caller's RDRAM
	FARM_NO_SCANOUT_COPY = 1u << 1, // diagnostic: scan out on the device only
};

struct farm_stream
{
	rdp_b200_processor *proc;
	const uint32_t *const *frames; // command words of the stream's distinct frames, cycled
	const uint64_t *frame_words;
	uint32_t num_frames;
	uint32_t depth;                // frames in flight (= number of output buffers)
	const uint32_t *vi_fields;     // 2 x 14 VI registers: even / odd field
	uint32_t *const *outs;         // `depth` pinned RGBA8 buffers
	uint64_t out_capacity;         // pixels per buffer
	rdp_b200_scanout_options opts;
	uint32_t expect_w, expect_h;   // 0: do not check
};

// One job of the pooled farm: a whole stream (its initial RDRAM image and `frames_to_run` frames) run on whichever
// processor of the matching kind becomes free.
struct farm_job
{
	uint32_t kind;                 // selects the processors that can run it (farm_slot::kind)
	const uint8_t *rdram_image;    // rdram_bytes of initial RDRAM (textures etc.); NULL: zeros
	const uint32_t *const *frames;
	const uint64_t *frame_words;
	uint32_t num_frames;           // distinct frames, cycled
	uint32_t frames_to_run;
	const uint32_t *vi_fields;
	uint32_t expect_w, expect_h;
	uint32_t downscale_steps;
};

struct farm_slot
{
	rdp_b200_processor *proc;
	uint32_t kind;
	uint32_t depth;
	uint32_t *const *outs;
	uint64_t out_capacity;
	uint8_t *rdram;                // the processor's borrowed host RDRAM
	uint64_t rdram_bytes;
};

}

// The driver binds to the rdp library it is told about (farm_bind) through dlsym on that library's own handle: it is
// loaded RTLD_LOCAL like any plugin, so its C++ symbols cannot interpose on another library's in the same process.
namespace
{
struct Api
{
	decltype(&rdp_b200_wait_for_timeline) wait_for_timeline;
	decltype(&rdp_b200_scanout_wait) scanout_wait;
	decltype(&rdp_b200_enqueue_command_stream_borrowed) enqueue_command_stream_borrowed;
	decltype(&rdp_b200_set_vi_registers) set_vi_registers;
	decltype(&rdp_b200_signal_timeline) signal_timeline;
	decltype(&rdp_b200_scanout) scanout;
	decltype(&rdp_b200_scanout_async) scanout_async;
	decltype(&rdp_b200_idle) idle;
	decltype(&rdp_b200_end_write_rdram) end_write_rdram;
	decltype(&rdp_b200_begin_read_hidden_rdram) begin_read_hidden_rdram;
	decltype(&rdp_b200_get_hidden_rdram_size) get_hidden_rdram_size;
	decltype(&rdp_b200_end_write_hidden_rdram) end_write_hidden_rdram;
	decltype(&rdp_b200_get_tmem) get_tmem;
	decltype(&rdp_b200_default_scanout_options) default_scanout_options;
} api;
}

extern "C" int farm_bind(const char *rdp_library_path)
{
	void *h = dlopen(rdp_library_path, RTLD_NOW | RTLD_LOCAL);
	if (!h)
	{
		fprintf(stderr, "farm_bind: %s\n", dlerror());
		return -1;
	}
	bool ok = true;
#define BIND(name) ok = ok && (api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "rdp_b200_" #name))) != nullptr
	BIND(wait_for_timeline); BIND(scanout_wait); BIND(enqueue_command_stream_borrowed); BIND(set_vi_registers);
	BIND(signal_timeline); BIND(scanout); BIND(scanout_async); BIND(idle); BIND(end_write_rdram);
	BIND(begin_read_hidden_rdram); BIND(get_hidden_rdram_size); BIND(end_write_hidden_rdram); BIND(get_tmem);
	BIND(default_scanout_options);
#undef BIND
	return ok ? 0 : -1;
}

namespace
{
struct InFlight
{
	uint64_t fence = 0;
	uint64_t timeline = 0;
	unsigned w = 0, h = 0;
};

struct StreamRunner
{
	const farm_stream *s = nullptr;
	std::vector<InFlight> ring;
	uint32_t issued = 0;

	bool retire(uint32_t slot, uint32_t flags)
	{
		InFlight &f = ring[slot];
		if (f.timeline)
		{
			api.wait_for_timeline(s->proc, f.timeline);
			f.timeline = 0;
		}
		if (f.fence)
		{
			api.scanout_wait(s->proc, f.fence);
			f.fence = 0;
			if (!(flags & FARM_NO_SCANOUT_COPY) && s->expect_w && (f.w != s->expect_w || f.h != s->expect_h))
				return false;
		}
		return true;
	}

	bool frame(uint32_t index, uint32_t flags)
	{
		const uint32_t slot = issued % s->depth;
		if (!retire(slot, flags))
			return false;
		const uint32_t fi = index % s->num_frames;
		api.enqueue_command_stream_borrowed(s->proc, size_t(s->frame_words[fi]), s->frames[fi]);
		api.set_vi_registers(s->proc, s->vi_fields + 14u * (index & 1u));
		InFlight &f = ring[slot];
		if (flags & FARM_READ_BACK_RDRAM)
			f.timeline = api.signal_timeline(s->proc);
		if (flags & FARM_NO_SCANOUT_COPY)
		{
			unsigned w = 0, h = 0;
			api.scanout(s->proc, &s->opts, &w, &h);
		}
		else
			f.fence = api.scanout_async(s->proc, &s->opts, s->outs[slot], size_t(s->out_capacity), &f.w, &f.h);
		issued++;
		return true;
	}

	bool finish(uint32_t flags)
	{
		bool ok = true;
		for (uint32_t i = 0; i < s->depth; i++)
			ok = retire(i, flags) && ok;
		return ok;
	}
};
}

extern "C" {

// Frames first_frame .. first_frame + frames_per_stream - 1 of every stream, round-robin over the streams a thread
// owns (stream i belongs to thread i % threads).  Returns 0, or -1 when a scan-out did not have the expected size.
int farm_run(const farm_stream *streams, uint32_t num_streams, uint32_t first_frame, uint32_t frames_per_stream,
             uint32_t threads, uint32_t flags)
{
	if (!num_streams || !frames_per_stream)
		return 0;
	threads = threads ? (threads < num_streams ? threads : num_streams) : 1u;
	std::atomic<int> failed{0};
	auto worker = [&](uint32_t t) {
		std::vector<StreamRunner> mine;
		for (uint32_t i = t; i < num_streams; i += threads)
		{
			StreamRunner r;
			r.s = &streams[i];
			r.ring.resize(streams[i].depth);
			mine.push_back(std::move(r));
		}
		for (uint32_t f = 0; f < frames_per_stream && !failed.load(std::memory_order_relaxed); f++)
			for (auto &r : mine)
				if (!r.frame(first_frame + f, flags))
					failed.store(1);
		for (auto &r : mine)
			if (!r.finish(flags))
				failed.store(1);
	};
	std::vector<std::thread> pool;
	for (uint32_t t = 1; t < threads; t++)
		pool.emplace_back(worker, t);
	worker(0);
	for (auto &th : pool)
		th.join();
	return failed.load() ? -1 : 0;
}

// Pooled farm: `num_jobs` whole streams over `num_slots` processors.  A slot runs one job at a time: reset (idle, the
// job's RDRAM image into the borrowed host RDRAM, end_write_rdram, hidden RDRAM and TMEM cleared), then the job's
// frames with `depth` in flight; a thread interleaves the slots it owns frame by frame.  Jobs are taken in order from
// one queue per kind.  frames_done receives the number of frames rendered.  Returns 0 / -1 as farm_run.
int farm_run_pool(const farm_slot *slots, uint32_t num_slots, const farm_job *jobs, uint32_t num_jobs, uint32_t threads,
                  uint32_t flags, uint64_t *frames_done)
{
	threads = threads ? (threads < num_slots ? threads : num_slots) : 1u;
	uint32_t max_kind = 0;
	for (uint32_t i = 0; i < num_jobs; i++)
		max_kind = jobs[i].kind > max_kind ? jobs[i].kind : max_kind;
	std::vector<std::vector<uint32_t>> queue(max_kind + 1u);
	for (uint32_t i = 0; i < num_jobs; i++)
		queue[jobs[i].kind].push_back(i);
	std::vector<std::atomic<uint32_t>> next(max_kind + 1u);
	for (auto &n : next)
		n.store(0);
	std::atomic<int> failed{0};
	std::atomic<uint64_t> total{0};

	struct SlotState
	{
		const farm_slot *slot = nullptr;
		farm_stream stream = {};
		StreamRunner runner;
		const farm_job *job = nullptr;
		uint32_t frame = 0;
		bool exhausted = false;
	};

	auto worker = [&](uint32_t t) {
		std::vector<SlotState> mine;
		for (uint32_t i = t; i < num_slots; i += threads)
		{
			SlotState st;
			st.slot = &slots[i];
			mine.push_back(st);
		}
		auto take_job = [&](SlotState &st) {
			const uint32_t kind = st.slot->kind;
			st.job = nullptr;
			if (kind > max_kind)
			{
				st.exhausted = true;
				return;
			}
			const uint32_t at = next[kind].fetch_add(1);
			if (at >= queue[kind].size())
			{
				st.exhausted = true;
				return;
			}
			const farm_job &job = jobs[queue[kind][at]];
			st.job = &job;
			st.frame = 0;
			rdp_b200_processor *p = st.slot->proc;
			api.idle(p);
			if (job.rdram_image)
				memcpy(st.slot->rdram, job.rdram_image, size_t(st.slot->rdram_bytes));
			else
				memset(st.slot->rdram, 0, size_t(st.slot->rdram_bytes));
			api.end_write_rdram(p);
			memset(api.begin_read_hidden_rdram(p), 0, api.get_hidden_rdram_size(p));
			api.end_write_hidden_rdram(p);
			memset(api.get_tmem(p), 0, 4096);
			farm_stream &s = st.stream;
			memset(&s, 0, sizeof(s));
			s.proc = p;
			s.frames = job.frames;
			s.frame_words = job.frame_words;
			s.num_frames = job.num_frames;
			s.depth = st.slot->depth;
			s.vi_fields = job.vi_fields;
			s.outs = st.slot->outs;
			s.out_capacity = st.slot->out_capacity;
			api.default_scanout_options(&s.opts);
			s.opts.blend_previous_frame = 1;
			s.opts.upscale_deinterlacing = 0;
			s.opts.downscale_steps = job.downscale_steps;
			s.expect_w = job.expect_w;
			s.expect_h = job.expect_h;
			st.runner = StreamRunner();
			st.runner.s = &st.stream;
			st.runner.ring.resize(s.depth);
		};
		for (auto &st : mine)
			take_job(st);
		for (;;)
		{
			bool any = false;
			for (auto &st : mine)
			{
				if (st.exhausted || failed.load(std::memory_order_relaxed))
					continue;
				any = true;
				if (!st.runner.frame(st.frame, flags))
					failed.store(1);
				total.fetch_add(1, std::memory_order_relaxed);
				if (++st.frame == st.job->frames_to_run)
				{
					if (!st.runner.finish(flags))
						failed.store(1);
					take_job(st);
				}
			}
			if (!any)
				break;
		}
	};
	std::vector<std::thread> pool;
	for (uint32_t t = 1; t < threads; t++)
		pool.emplace_back(worker, t);
	worker(0);
	for (auto &th : pool)
		th.join();
	if (frames_done)
		*frames_done = total.load();
	return failed.load() ? -1 : 0;
}

}
