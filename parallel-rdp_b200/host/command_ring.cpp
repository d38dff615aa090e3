// command_ring.cpp -- see command_ring.hpp.
#include "command_ring.hpp"
#include "rdp_device.hpp"
#include "rdp_setup.cuh"
#include <chrono>
#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

namespace RDP
{
namespace
{
inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
	__builtin_ia32_pause();
#elif defined(__aarch64__)
	asm volatile("yield" ::: "memory");
#endif
}

// Polls `pred` for at most `budget_ns` without giving up the core; true when it came true.
template <class Pred>
inline bool spin_for(int64_t budget_ns, Pred pred)
{
	if (pred())
		return true;
	if (budget_ns <= 0)
		return false;
	const auto t0 = std::chrono::steady_clock::now();
	for (;;)
	{
		for (int i = 0; i < 32; i++)
		{
			cpu_relax();
			if (pred())
				return true;
		}
		if (std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() >= budget_ns)
			return false;
	}
}
}

void CommandRing::init(CommandProcessor *processor_, unsigned count_words)
{
	teardown_thread();
	processor = processor_;
	ring.assign(count_words, 0);
	size = count_words;
	write_count = 0;
	read_count = 0;
	producer_pos = 0;
	if (const char *env = getenv("RDP_B200_RING_SPIN_US"))
		spin_ns = int64_t(strtol(env, nullptr, 0)) * 1000;
	thr = std::thread(&CommandRing::thread_loop, this);
}

void CommandRing::set_affinity(const std::vector<int> &cpus)
{
#if defined(__linux__)
	if (cpus.empty() || !thr.joinable())
		return;
	cpu_set_t set;
	CPU_ZERO(&set);
	for (int c : cpus)
		if (c >= 0 && c < CPU_SETSIZE)
			CPU_SET(c, &set);
	pthread_setaffinity_np(thr.native_handle(), sizeof(set), &set);
#else
	(void)cpus;
#endif
}

void CommandRing::teardown_thread()
{
	if (thr.joinable())
	{
		reserve(0, kHeaderExit);
		publish(0);
		thr.join();
	}
}

CommandRing::~CommandRing()
{
	teardown_thread();
}

// The producer waits for the consumer: a short spin, then sleep until the consumer (which checks producer_asleep
// after every record) signals.  The timeout only bounds the cost of a wake-up lost to a spurious ordering.
template <class Pred>
void CommandRing::producer_wait(Pred pred)
{
	if (spin_for(spin_ns, pred))
		return;
	std::unique_lock<std::mutex> holder{lock};
	while (!pred())
	{
		producer_asleep.store(true, std::memory_order_seq_cst);
		if (pred())
			break;
		cond_producer.wait_for(holder, std::chrono::milliseconds(2));
	}
	producer_asleep.store(false, std::memory_order_relaxed);
}

// Returns a contiguous payload area of `payload_words` words preceded by `header` (already written).
uint32_t *CommandRing::reserve(uint32_t payload_words, uint32_t header)
{
	const uint64_t need = uint64_t(payload_words) + 1;
	uint64_t offset = producer_pos % size;
	uint64_t total = need;
	const bool pad = offset + need > size;
	if (pad)
		total += size - offset; // the tail is skipped

	// Wait for space (rare: the ring holds several frames' worth of commands).
	producer_wait([&]() { return producer_pos + total <= read_count.load(std::memory_order_acquire) + size; });

	if (pad)
	{
		ring[offset] = kHeaderPad;
		producer_pos += size - offset;
		offset = 0;
	}
	ring[offset] = header;
	return ring.data() + offset + 1;
}

void CommandRing::publish(uint32_t payload_words)
{
	producer_pos += uint64_t(payload_words) + 1;
	write_count.store(producer_pos, std::memory_order_release);
	std::atomic_thread_fence(std::memory_order_seq_cst);
	if (consumer_asleep.load(std::memory_order_relaxed))
	{
		std::lock_guard<std::mutex> holder{lock};
		cond_consumer.notify_one();
	}
}

void CommandRing::enqueue_command(unsigned num_words, const uint32_t *words)
{
	if (!num_words)
		return;
	uint32_t *dst = reserve(num_words, num_words);
	memcpy(dst, words, size_t(num_words) * sizeof(uint32_t));
	publish(num_words);
}

void CommandRing::enqueue_meta(RingMeta kind, unsigned num_words, const uint32_t *payload)
{
	uint32_t *dst = reserve(num_words + 1u, (num_words + 1u) | kHeaderMeta);
	dst[0] = uint32_t(kind);
	if (num_words)
		memcpy(dst + 1, payload, size_t(num_words) * sizeof(uint32_t));
	publish(num_words + 1u);
}

void CommandRing::enqueue_stream(size_t num_words, const uint32_t *words)
{
	size_t pos = 0;
	while (pos < num_words)
	{
		// Cut at a command boundary at or below kMaxRecordWords.
		size_t end = pos;
		while (end < num_words)
		{
			const size_t len = rdpb::command_length_words((words[end] >> 24) & 63u);
			if (end + len > num_words || end + len - pos > kMaxRecordWords)
				break;
			end += len;
		}
		if (end == pos)
			break; // truncated trailing command: dropped, as a short enqueue_command would be
		const uint32_t n = uint32_t(end - pos);
		uint32_t *dst = reserve(n, n | kHeaderStream);
		memcpy(dst, words + pos, size_t(n) * sizeof(uint32_t));
		publish(n);
		pos = end;
	}
}

void CommandRing::enqueue_stream_borrowed(size_t num_words, const uint32_t *words)
{
	if (!num_words)
		return;
	uint32_t *dst = reserve(4, 4u | kHeaderBorrowed);
	const uint64_t ptr = uint64_t(reinterpret_cast<uintptr_t>(words)), n = uint64_t(num_words);
	dst[0] = uint32_t(ptr);
	dst[1] = uint32_t(ptr >> 32);
	dst[2] = uint32_t(n);
	dst[3] = uint32_t(n >> 32);
	publish(4);
}

void CommandRing::drain()
{
	const uint64_t target = producer_pos;
	producer_wait([&]() { return read_count.load(std::memory_order_acquire) >= target; });
}

void CommandRing::wait_until_at_least(const std::atomic<uint64_t> &flag, uint64_t value)
{
	producer_wait([&]() { return flag.load(std::memory_order_acquire) >= value; });
}

void CommandRing::thread_loop()
{
	uint64_t r = 0;
	bool idle_kick_due = false; // records were executed since the last idle kick
	for (;;)
	{
		// Wait for a record: a short spin, then sleep.  500 us after the last record the processor is told, once, that
		// the producer went quiet (it kicks whatever it has batched); after that the thread sleeps until a publish.
		const auto have_record = [&]() { return write_count.load(std::memory_order_acquire) != r; };
		if (!spin_for(spin_ns, have_record))
		{
			bool idle = false;
			{
				std::unique_lock<std::mutex> holder{lock};
				consumer_asleep.store(true, std::memory_order_seq_cst);
				std::atomic_thread_fence(std::memory_order_seq_cst);
				if (idle_kick_due)
					idle = !cond_consumer.wait_for(holder, std::chrono::microseconds(500), have_record);
				else
					cond_consumer.wait(holder, have_record);
				consumer_asleep.store(false, std::memory_order_relaxed);
			}
			if (idle)
			{
				idle_kick_due = false;
				processor->execute_meta(RingMeta::Idle, nullptr, 0);
				continue;
			}
		}

		uint64_t offset = r % size;
		uint32_t header = ring[offset];
		if (header == kHeaderPad)
		{
			r += size - offset;
			offset = 0;
			header = ring[0];
		}
		if (header == kHeaderExit)
		{
			read_count.store(r + 1, std::memory_order_release);
			break;
		}

		const auto busy_t0 = std::chrono::steady_clock::now();
		const uint32_t n = header & ~uint32_t(kHeaderFlags);
		const uint32_t *payload = ring.data() + offset + 1;
		if (header & kHeaderMeta)
			processor->execute_meta(RingMeta(payload[0]), payload + 1, n - 1u);
		else if (header & kHeaderBorrowed)
		{
			const uint32_t *words = reinterpret_cast<const uint32_t *>(uintptr_t(uint64_t(payload[0]) | (uint64_t(payload[1]) << 32)));
			const uint64_t total = uint64_t(payload[2]) | (uint64_t(payload[3]) << 32);
			processor->execute_borrowed_stream(size_t(total), words);
		}
		else if (header & kHeaderStream)
			processor->execute_command_run(n, payload);
		else
			processor->enqueue_command_direct(n, payload);
		idle_kick_due = true;

		processor->add_ring_busy_ns(uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - busy_t0).count()));
		// The record was executed in place: only now may the producer reuse its words.
		r += uint64_t(n) + 1;
		read_count.store(r, std::memory_order_release);
		std::atomic_thread_fence(std::memory_order_seq_cst);
		if (producer_asleep.load(std::memory_order_relaxed))
		{
			std::lock_guard<std::mutex> holder{lock};
			cond_producer.notify_all();
		}
	}
}
}
