// A small process-wide pool of helper threads for the host side of the greedy loop.
//
// The decisions of the reads BEHIND the commit frontier of a chunk are independent of each other (they are evaluated
// against a frozen set of founders, only to find out which device work to request next), and the expensive part of one —
// reproducing the reference's candidate order for a read with tied candidates (ref_order.h) — is tens of microseconds of
// scalar work.  A batch worker hands those evaluations to this pool and takes part itself; the pool matters most for the
// last batches of a run, when most workers are done and their cores idle.  Results never depend on the pool size:
// every task writes only its own positions.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

// Host cores this process may run on: the affinity mask (taskset, cgroup cpusets, container limits), not the machine's
// core count; ICB_HOST_CORES overrides it (a launcher that runs one process per GPU passes each rank its share).
inline int icb_host_cores()
{
    if (const char* e = getenv("ICB_HOST_CORES")) {
	const int v = atoi(e);
	if (v > 0) return v;
    }
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
	const int c = CPU_COUNT(&set);
	if (c > 0) return c;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? int(h) : 1;
}

class IcbHelperPool {
public:
    static IcbHelperPool& get()
    {
	static IcbHelperPool pool;
	return pool;
    }
    int helpers() const { return int(threads_.size()); }
    // fn(i) for every i in [0, n): on the helpers and on the calling thread; returns when all of them have finished
    void parallel_for(int n, const std::function<void(int)>& fn)
    {
	if (n <= 0) return;
	if (n == 1 || threads_.empty()) {
	    for (int i = 0; i < n; i++) fn(i);
	    return;
	}
	auto job = std::make_shared<Job>();
	job->fn = &fn;
	job->n = n;
	{
	    std::lock_guard<std::mutex> g(m_);
	    jobs_.push_back(job);
	}
	cv_.notify_all();
	for (;;) {
	    const int i = job->next.fetch_add(1);
	    if (i >= n) break;
	    fn(i);
	    job->done.fetch_add(1);
	}
	{
	    std::lock_guard<std::mutex> g(m_);
	    for (auto it = jobs_.begin(); it != jobs_.end(); ++it)
		if (it->get() == job.get()) {
		    jobs_.erase(it);
		    break;
		}
	}
	while (job->done.load(std::memory_order_acquire) < n) std::this_thread::yield();
    }

private:
    struct Job {
	const std::function<void(int)>* fn = nullptr;
	int n = 0;
	std::atomic<int> next{0}, done{0};
    };
    IcbHelperPool()
    {
	int want = icb_host_cores();
	if (want > 16) want = 16;
	if (const char* e = getenv("ICB_HELPERS")) want = atoi(e);
	for (int t = 0; t < want; t++) threads_.emplace_back([this] { run(); });
    }
    ~IcbHelperPool()
    {
	{
	    std::lock_guard<std::mutex> g(m_);
	    stop_ = true;
	}
	cv_.notify_all();
	for (auto& t : threads_) t.join();
    }
    void run()
    {
	std::unique_lock<std::mutex> lk(m_);
	for (;;) {
	    cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
	    if (stop_) return;
	    std::shared_ptr<Job> job = jobs_.front();
	    const int i = job->next.fetch_add(1);
	    if (i >= job->n) {
		if (!jobs_.empty() && jobs_.front().get() == job.get()) jobs_.pop_front();
		continue;
	    }
	    lk.unlock();
	    (*job->fn)(i);
	    job->done.fetch_add(1, std::memory_order_release);
	    lk.lock();
	}
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> jobs_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};
