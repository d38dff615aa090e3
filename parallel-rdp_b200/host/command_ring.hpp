// command_ring.hpp -- single-producer / single-consumer ring between the emulator thread and the command thread.
//
// Role mirrors parallel-rdp/command_ring.{hpp,cpp} (4096-word ring, worker thread, 500 us idle => MetaIdle).
// The reference takes a mutex and signals a condition variable on every command; at the 1e8 commands/s this
// project targets that alone exceeds the per-command budget, so this ring is lock-free: the producer appends
// [count, words...] records and publishes the write index with one release store; the mutex/condvar pair is only
// touched when one side actually has to sleep.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <stdint.h>
#include <thread>
#include <vector>

namespace RDP
{
class CommandProcessor;

class CommandRing
{
public:
	~CommandRing();
	void init(CommandProcessor *processor, unsigned count_words);
	bool is_running() const { return thr.joinable(); }
	void enqueue_command(unsigned num_words, const uint32_t *words);
	void drain();
	void shutdown() { teardown_thread(); }

private:
	CommandProcessor *processor = nullptr;
	std::thread thr;
	std::vector<uint32_t> ring;
	uint64_t mask = 0;

	alignas(64) std::atomic<uint64_t> write_count{0};     // published by the producer
	alignas(64) std::atomic<uint64_t> read_count{0};      // space released by the consumer
	alignas(64) std::atomic<uint64_t> completed_count{0}; // commands fully executed up to here
	std::atomic<bool> consumer_asleep{false};
	std::atomic<bool> producer_asleep{false};
	std::atomic<bool> dead{false};

	std::mutex lock;
	std::condition_variable cond_consumer;
	std::condition_variable cond_producer;

	void thread_loop();
	void teardown_thread();
};
}
