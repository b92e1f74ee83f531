// TEST INFRASTRUCTURE (oracle).  `hls::stream<T>` as an unbounded FIFO.  In C
// simulation the reference's DATAFLOW regions run sequentially and every
// producer is called before its consumer (all six */src/conv_layer.cc), so an
// unbounded queue reproduces the hardware FIFOs' contents exactly.
#ifndef FLOWGNN_ORACLE_SHIM_HLS_STREAM_H
#define FLOWGNN_ORACLE_SHIM_HLS_STREAM_H

#include <cstdio>
#include <cstdlib>
#include <deque>

namespace hls {
template <typename T>
class stream
{
public:
    stream() {}
    explicit stream(const char*) {}
    stream(const stream&) = delete;
    stream& operator=(const stream&) = delete;

    void operator<<(const T& x) { q_.push_back(x); }
    void operator>>(T& x)
    {
        if (q_.empty())
        {
            std::fprintf(stderr, "oracle shim: read from empty hls::stream\n");
            std::abort();
        }
        x = q_.front();
        q_.pop_front();
    }
    bool empty() const { return q_.empty(); }

private:
    std::deque<T> q_;
};
}

#endif
