// command_ring.cpp -- see command_ring.hpp.
#include "command_ring.hpp"
#include "rdp_device.hpp"
#include "rdp_setup.cuh"
#include <chrono>
#include <string.h>

namespace RDP
{
void CommandRing::init(CommandProcessor *processor_, unsigned count_words)
{
	teardown_thread();
	processor = processor_;
	ring.assign(count_words, 0);
	size = count_words;
	write_count = 0;
	read_count = 0;
	producer_pos = 0;
	thr = std::thread(&CommandRing::thread_loop, this);
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
	while (producer_pos + total > read_count.load(std::memory_order_acquire) + size)
	{
		std::unique_lock<std::mutex> holder{lock};
		producer_asleep.store(true, std::memory_order_seq_cst);
		cond_producer.wait_for(holder, std::chrono::microseconds(100), [&]() {
			return producer_pos + total <= read_count.load(std::memory_order_acquire) + size;
		});
		producer_asleep.store(false, std::memory_order_relaxed);
	}

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
	unsigned spins = 0;
	while (read_count.load(std::memory_order_acquire) < target)
	{
		if (++spins < 2000)
			std::this_thread::yield();
		else
			std::this_thread::sleep_for(std::chrono::microseconds(20));
	}
}

void CommandRing::thread_loop()
{
	uint64_t r = 0;
	for (;;)
	{
		// Wait for a record; after 500 us without traffic tell the renderer the producer went idle.
		bool idle = false;
		unsigned spins = 0;
		while (write_count.load(std::memory_order_acquire) == r)
		{
			if (++spins < 4000)
			{
				std::this_thread::yield();
				continue;
			}
			std::unique_lock<std::mutex> holder{lock};
			consumer_asleep.store(true, std::memory_order_seq_cst);
			std::atomic_thread_fence(std::memory_order_seq_cst);
			bool got = cond_consumer.wait_for(holder, std::chrono::microseconds(500), [&]() {
				return write_count.load(std::memory_order_acquire) != r;
			});
			consumer_asleep.store(false, std::memory_order_relaxed);
			if (!got)
			{
				idle = true;
				break;
			}
		}

		if (idle)
		{
			const uint32_t idle_word = 3u << 24; // Op::MetaIdle
			processor->enqueue_command_direct(1, &idle_word);
			continue;
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
		const uint32_t n = header & ~uint32_t(kHeaderStream | kHeaderBorrowed);
		const uint32_t *payload = ring.data() + offset + 1;
		if (header & kHeaderBorrowed)
		{
			const uint32_t *words = reinterpret_cast<const uint32_t *>(uintptr_t(uint64_t(payload[0]) | (uint64_t(payload[1]) << 32)));
			const uint64_t total = uint64_t(payload[2]) | (uint64_t(payload[3]) << 32);
			processor->execute_borrowed_stream(size_t(total), words);
		}
		else if (header & kHeaderStream)
		{
			uint32_t pos = 0;
			while (pos < n)
			{
				const uint32_t len = rdpb::command_length_words((payload[pos] >> 24) & 63u);
				processor->enqueue_command_direct(len, payload + pos);
				pos += len;
			}
		}
		else
			processor->enqueue_command_direct(n, payload);

		processor->add_ring_busy_ns(uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - busy_t0).count()));
		// The record was executed in place: only now may the producer reuse its words.
		r += uint64_t(n) + 1;
		read_count.store(r, std::memory_order_release);
		if (producer_asleep.load(std::memory_order_seq_cst))
		{
			std::lock_guard<std::mutex> holder{lock};
			cond_producer.notify_one();
		}
	}
}
}
