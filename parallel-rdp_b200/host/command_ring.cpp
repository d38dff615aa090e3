// command_ring.cpp -- see command_ring.hpp.
#include "command_ring.hpp"
#include "rdp_device.hpp"
#include <chrono>

namespace RDP
{
void CommandRing::init(CommandProcessor *processor_, unsigned count_words)
{
	teardown_thread();
	processor = processor_;
	ring.assign(count_words, 0);
	mask = count_words - 1;
	write_count = 0;
	read_count = 0;
	completed_count = 0;
	dead = false;
	thr = std::thread(&CommandRing::thread_loop, this);
}

void CommandRing::teardown_thread()
{
	if (thr.joinable())
	{
		enqueue_command(0, nullptr); // zero-length record = sentinel
		thr.join();
	}
}

CommandRing::~CommandRing()
{
	teardown_thread();
}

void CommandRing::enqueue_command(unsigned num_words, const uint32_t *words)
{
	const uint64_t need = uint64_t(num_words) + 1;
	uint64_t w = write_count.load(std::memory_order_relaxed);

	// Wait for space (rare: the ring is far larger than a frame's worth of state commands).
	while (w + need > read_count.load(std::memory_order_acquire) + ring.size())
	{
		std::unique_lock<std::mutex> holder{lock};
		producer_asleep.store(true, std::memory_order_seq_cst);
		cond_producer.wait_for(holder, std::chrono::microseconds(100), [&]() {
			return w + need <= read_count.load(std::memory_order_acquire) + ring.size();
		});
		producer_asleep.store(false, std::memory_order_relaxed);
	}

	ring[w++ & mask] = num_words;
	for (unsigned i = 0; i < num_words; i++)
		ring[w++ & mask] = words[i];
	write_count.store(w, std::memory_order_seq_cst);

	if (consumer_asleep.load(std::memory_order_seq_cst))
	{
		std::lock_guard<std::mutex> holder{lock};
		cond_consumer.notify_one();
	}
}

void CommandRing::drain()
{
	const uint64_t target = write_count.load(std::memory_order_relaxed);
	unsigned spins = 0;
	while (completed_count.load(std::memory_order_acquire) < target)
	{
		if (++spins < 2000)
			std::this_thread::yield();
		else
			std::this_thread::sleep_for(std::chrono::microseconds(20));
	}
}

void CommandRing::thread_loop()
{
	std::vector<uint32_t> cmd;
	cmd.reserve(64);
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

		const uint32_t num_words = ring[r++ & mask];
		cmd.resize(num_words);
		for (uint32_t i = 0; i < num_words; i++)
			cmd[i] = ring[r++ & mask];
		read_count.store(r, std::memory_order_release);
		if (producer_asleep.load(std::memory_order_seq_cst))
		{
			std::lock_guard<std::mutex> holder{lock};
			cond_producer.notify_one();
		}

		if (cmd.empty())
			break;

		processor->enqueue_command_direct(num_words, cmd.data());
		completed_count.store(r, std::memory_order_release);
	}
	completed_count.store(r, std::memory_order_release);
}
}
