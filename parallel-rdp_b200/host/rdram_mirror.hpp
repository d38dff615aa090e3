// rdram_mirror.hpp -- bookkeeping for the two-way RDRAM mirror between the caller's host range and device RDRAM.
//
// Replaces VK_EXT_external_memory_host / the reference's "incoherent" path (parallel-rdp/rdp_renderer.cpp:2523-2907:
// 1 KiB pages, page_to_direct_copy / page_to_masked_copy / page_to_pending_readback bitmaps, and the masked copy-back
// on the fence thread, rdp_device.cpp:1195-1258).  Same granularity -- 1 KiB pages -- different contract:
//
//   host -> device   The CPU side says what it wrote (CommandProcessor::mark_rdram_written, or end_write_rdram for
//                    "anything may have changed").  Whole pages inside the range go into a dirty-page bitmap; the
//                    partial first / last page is kept as an exact byte range, so an upload never carries bytes the
//                    CPU did not declare -- in particular never stale host bytes over pixels the GPU has rendered
//                    into the same page and not yet returned ("GPU wins", as in the reference).
//   device -> host   Every pass notes the exact byte ranges it renders to (colour image, depth image).  A timeline
//                    signal copies those ranges -- not a bounding box, so CPU data lying between the colour and the
//                    depth image is never touched -- back on the read-back stream.
//
// mark_* may be called from the emulator thread while the command thread consumes: a mutex guards the lists, an
// atomic flag keeps the common "nothing dirty" check free.
#pragma once
#include <algorithm>
#include <atomic>
#include <mutex>
#include <stdint.h>
#include <utility>
#include <vector>

namespace RDP
{
class RdramMirror
{
public:
	enum : uint32_t { kPageShift = 10, kPageSize = 1u << kPageShift, kMaxFragments = 1024, kMaxGpuRanges = 32 };

	struct Stats
	{
		uint64_t h2d_bytes = 0, h2d_copies = 0, d2h_bytes = 0, d2h_copies = 0;
	};

	void init(size_t rdram_size_)
	{
		rdram_size = uint32_t(rdram_size_);
		dirty_pages.assign((rdram_size / kPageSize + 63u) / 64u, 0ull);
		fragments.clear();
		gpu_ranges.clear();
		all_dirty = false;
		any_dirty.store(false, std::memory_order_relaxed);
	}

	// ---- CPU writes ----
	void mark_all_written()
	{
		std::lock_guard<std::mutex> holder{lock};
		all_dirty = true;
		any_dirty.store(true, std::memory_order_release);
	}

	void mark_written(uint64_t offset, uint64_t size)
	{
		if (!size || offset >= rdram_size)
			return;
		const uint64_t end = std::min<uint64_t>(offset + size, rdram_size);
		std::lock_guard<std::mutex> holder{lock};
		if (all_dirty)
			return;
		const uint64_t first_full = (offset + kPageSize - 1) >> kPageShift, last_full = end >> kPageShift; // [first, last)
		if (first_full < last_full)
		{
			for (uint64_t p = first_full; p < last_full; p++)
				dirty_pages[p >> 6] |= 1ull << (p & 63u);
			add_fragment(uint32_t(offset), uint32_t(first_full << kPageShift));
			add_fragment(uint32_t(last_full << kPageShift), uint32_t(end));
		}
		else
			add_fragment(uint32_t(offset), uint32_t(end)); // inside one page, or straddling one page boundary
		any_dirty.store(true, std::memory_order_release);
	}

	bool has_host_writes() const { return any_dirty.load(std::memory_order_acquire); }
	bool whole_range_dirty()
	{
		std::lock_guard<std::mutex> holder{lock};
		return all_dirty;
	}

	// Hands every pending host write to `copy(offset, size)` as maximal contiguous runs and forgets them.
	template <class F>
	void consume_host_writes(F &&copy)
	{
		std::vector<std::pair<uint32_t, uint32_t>> runs;
		{
			std::lock_guard<std::mutex> holder{lock};
			any_dirty.store(false, std::memory_order_relaxed);
			if (all_dirty)
				runs.emplace_back(0u, rdram_size);
			else
			{
				for (auto &f : fragments)
					runs.emplace_back(f.first, f.second - f.first);
				const uint32_t num_pages = rdram_size >> kPageShift;
				uint32_t p = 0;
				while (p < num_pages)
				{
					const uint64_t word = dirty_pages[p >> 6] >> (p & 63u);
					if (!word)
					{
						p = (p | 63u) + 1u;
						continue;
					}
					p += uint32_t(__builtin_ctzll(word));
					uint32_t q = p;
					while (q < num_pages && ((dirty_pages[q >> 6] >> (q & 63u)) & 1ull))
						q++;
					runs.emplace_back(p << kPageShift, (q - p) << kPageShift);
					p = q;
				}
			}
			all_dirty = false;
			fragments.clear();
			std::fill(dirty_pages.begin(), dirty_pages.end(), 0ull);
		}
		// Merge touching runs (a fragment next to the pages it belongs to) so the caller issues fewer copies.
		std::sort(runs.begin(), runs.end());
		size_t out = 0;
		for (size_t i = 0; i < runs.size(); i++)
		{
			if (out && runs[out - 1].first + runs[out - 1].second >= runs[i].first)
			{
				const uint32_t hi = std::max(runs[out - 1].first + runs[out - 1].second, runs[i].first + runs[i].second);
				runs[out - 1].second = hi - runs[out - 1].first;
			}
			else
				runs[out++] = runs[i];
		}
		for (size_t i = 0; i < out; i++)
		{
			stats.h2d_bytes += runs[i].second;
			stats.h2d_copies++;
			copy(runs[i].first, runs[i].second);
		}
	}

	// ---- GPU writes (command thread only) ----
	void note_gpu_write(uint32_t addr, uint32_t bytes)
	{
		if (!bytes)
			return;
		addr &= rdram_size - 1u;
		if (uint64_t(addr) + bytes > rdram_size)
		{
			// Wraps around the end of RDRAM: two ranges.
			const uint32_t head = rdram_size - addr;
			add_gpu_range(addr, rdram_size);
			add_gpu_range(0, std::min(bytes - head, rdram_size));
			return;
		}
		add_gpu_range(addr, addr + bytes);
	}

	bool has_gpu_writes() const { return !gpu_ranges.empty(); }
	const std::vector<std::pair<uint32_t, uint32_t>> &pending_gpu_ranges() const { return gpu_ranges; } // [lo, hi)

	template <class F>
	void consume_gpu_writes(F &&copy)
	{
		for (auto &r : gpu_ranges)
		{
			stats.d2h_bytes += r.second - r.first;
			stats.d2h_copies++;
			copy(r.first, r.second - r.first);
		}
		gpu_ranges.clear();
	}

	// ---- read-backs in flight (command thread only): device bytes newer than the host's until the copy has landed ----
	void add_inflight(uint32_t lo, uint32_t hi, uint64_t fence) { inflight.push_back({ lo, hi, fence }); }
	// `complete(fence)` -> true once the read-back behind that fence has landed; forgets those.
	template <class F>
	void retire_inflight(F &&complete)
	{
		size_t out = 0;
		for (size_t i = 0; i < inflight.size(); i++)
			if (!complete(inflight[i].fence))
				inflight[out++] = inflight[i];
		inflight.resize(out);
	}
	bool has_inflight() const { return !inflight.empty(); }

	// The parts of [lo, hi) for which host memory is authoritative: everything except what passes have rendered to and
	// not yet returned (pending, or a read-back still in flight).  The reference's rule -- "writes made by the GPU which
	// are not known to be resolved ... will always win over writes made by CPU" (rdp_renderer.cpp:2709-2716) -- at range
	// instead of byte granularity.
	void host_authoritative_runs(uint32_t lo, uint32_t hi, std::vector<std::pair<uint32_t, uint32_t>> &out) const
	{
		std::vector<std::pair<uint32_t, uint32_t>> owned;
		for (auto &r : gpu_ranges)
			if (r.first < hi && lo < r.second)
				owned.emplace_back(std::max(r.first, lo), std::min(r.second, hi));
		for (auto &r : inflight)
			if (r.lo < hi && lo < r.hi)
				owned.emplace_back(std::max(r.lo, lo), std::min(r.hi, hi));
		std::sort(owned.begin(), owned.end());
		uint32_t at = lo;
		for (auto &o : owned)
		{
			if (o.first > at)
				out.emplace_back(at, o.first - at);
			at = std::max(at, o.second);
		}
		if (at < hi)
			out.emplace_back(at, hi - at);
	}
	void count_upload(uint32_t bytes)
	{
		stats.h2d_bytes += bytes;
		stats.h2d_copies++;
	}

	Stats get_stats() const { return stats; }

private:
	uint32_t rdram_size = 0;
	std::mutex lock;
	std::atomic<bool> any_dirty{false};
	bool all_dirty = false;
	std::vector<uint64_t> dirty_pages;
	std::vector<std::pair<uint32_t, uint32_t>> fragments; // [lo, hi) byte ranges smaller than a page
	std::vector<std::pair<uint32_t, uint32_t>> gpu_ranges; // sorted, disjoint
	struct Inflight { uint32_t lo, hi; uint64_t fence; };
	std::vector<Inflight> inflight;
	Stats stats;

	void add_fragment(uint32_t lo, uint32_t hi)
	{
		if (lo >= hi)
			return;
		if (fragments.size() >= kMaxFragments)
		{
			// Too many distinct partial pages: fall back to whole pages for this one.
			for (uint32_t p = lo >> kPageShift; p <= (hi - 1u) >> kPageShift; p++)
				dirty_pages[p >> 6] |= 1ull << (p & 63u);
			return;
		}
		if (!fragments.empty() && fragments.back().second == lo)
			fragments.back().second = hi; // sequential small writes grow one fragment
		else
			fragments.emplace_back(lo, hi);
	}

	void add_gpu_range(uint32_t lo, uint32_t hi)
	{
		// Insert and merge with every range it touches or overlaps.
		size_t i = 0;
		while (i < gpu_ranges.size())
		{
			auto &r = gpu_ranges[i];
			if (r.second < lo || hi < r.first)
			{
				i++;
				continue;
			}
			lo = std::min(lo, r.first);
			hi = std::max(hi, r.second);
			gpu_ranges.erase(gpu_ranges.begin() + long(i));
		}
		gpu_ranges.emplace_back(lo, hi);
		std::sort(gpu_ranges.begin(), gpu_ranges.end());
		if (gpu_ranges.size() > kMaxGpuRanges)
		{
			// Join the two closest neighbours (copies a gap of device bytes the GPU did not write: still device truth).
			size_t best = 0;
			uint32_t best_gap = ~0u;
			for (size_t k = 0; k + 1 < gpu_ranges.size(); k++)
			{
				const uint32_t gap = gpu_ranges[k + 1].first - gpu_ranges[k].second;
				if (gap < best_gap)
				{
					best_gap = gap;
					best = k;
				}
			}
			gpu_ranges[best].second = gpu_ranges[best + 1].second;
			gpu_ranges.erase(gpu_ranges.begin() + long(best) + 1);
		}
	}
};
}
