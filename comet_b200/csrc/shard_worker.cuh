// shard_worker.cuh -- host-side helper of the single-process multi-GPU indexes (sharded.cu, ivf.cu): one persistent
// thread per shard enqueues that shard's part of a search, so that W devices start together.
#pragma once

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>

#include "common.cuh"

// One persistent thread per shard: takes a job, runs it with the shard's device current, hands back status + message
// (cm::fail writes a thread-local string, so the text has to travel with the status).
struct ShardWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, stop = false;
    int rc = CM_OK;
    std::string msg;
    explicit ShardWorker(int device) {
        th = std::thread([this, device] {
            cudaSetDevice(device);
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [&] { return has_job || stop; });
                if (stop) return;
                std::function<int()> f = std::move(job);
                has_job = false;
                lk.unlock();
                int r = f();
                std::string m = r == CM_OK ? std::string() : std::string(cm_last_error());
                lk.lock();
                rc = r; msg.swap(m); done = true;
                cv.notify_all();
            }
        });
    }
    ~ShardWorker() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
    void post(std::function<int()> f) {
        { std::lock_guard<std::mutex> lk(mu); job = std::move(f); has_job = true; done = false; }
        cv.notify_all();
    }
    int wait(std::string *m) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        if (m) *m = msg;
        return rc;
    }
};

