/* Shim: Dune::Timer (dune-common is not in this image). */
#ifndef B200_REF_SHIM_DUNE_TIMER_HH
#define B200_REF_SHIM_DUNE_TIMER_HH
#include <chrono>
namespace Dune {
class Timer {
    using clock = std::chrono::steady_clock;
    clock::time_point t0_; double acc_ = 0.0; bool running_;
public:
    explicit Timer(bool start = true) : t0_(clock::now()), running_(start) {}
    void reset() { acc_ = 0.0; t0_ = clock::now(); }
    void start() { if (!running_) { running_ = true; t0_ = clock::now(); } }
    double elapsed() const { return acc_ + (running_ ? std::chrono::duration<double>(clock::now() - t0_).count() : 0.0); }
    double stop() { if (running_) { acc_ += std::chrono::duration<double>(clock::now() - t0_).count(); running_ = false; } return acc_; }
};
}
#endif
