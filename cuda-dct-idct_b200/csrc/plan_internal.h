// plan_internal.h -- the plan object behind the opaque b200dct_plan handle (host TUs only).
#pragma once
#include "b200dct.h"
#include "any_kernels.cuh"
#include "rgb_kernels.cuh"

struct b200dct_plan {
    float T[64];
    float Q[64];
    float Qc[64];    // chrominance table of the colour entry point (default: ITU-T T.81 Annex K.2)
    uint64_t mask;
    bool sparse;     // T is bit-identical to Haweel's matrix
    bool q_default;  // Q is the JPEG luminance table
    bool q_fastdiv;  // every divisor is in the exhaustively proven set (integers 1..255)
    int path;        // b200dct_path
    int inverse;     // b200dct_inverse_mode
    int dense;       // b200dct_dense_mode
    bool symmetric;  // dense T whose even rows are symmetric and odd rows antisymmetric
    int tk;          // TK_HAWEEL / TK_DENSE_SYM / TK_DENSE: the kernels this plan runs
    b200dct::CommonParams cp; // device-ready tables
    b200dct::QuantTables qc; // device-ready chrominance tables (same mask)
    bool qc_default, qc_fastdiv;
};


namespace b200dct {
// Early path of a hardware-scheduled (direct-style) kernel across a launch boundary, see "early loads" in
// b200dct.cu: the scope holds the per-stream launch-order lock from the decision until done().
struct EarlyParams {
    int early = 0;                       // leading CTAs that may load and compute before griddepcontrol.wait
    int chain_feed = 0;                  // trailing CTAs that add 1 to the completion counter at their end
    unsigned long long *chain = nullptr; // completion counter of the (device, stream) record
    unsigned long long chain_target = 0; // its value once the predecessor is complete
};
struct Span { const void *ptr; size_t pitch, row_bytes; int rows; };
class EarlyScope {
public:
    // usable: programmatic dependent launch is on for this launch and the stream is not capturing;
    // ctas_per_sm: occupancy of the kernel about to be launched; rd: the plane it reads (ptr NULL: never early)
    EarlyScope(cudaStream_t s, bool usable, unsigned long long ctas, int ctas_per_sm, Span rd, Span w0, Span w1);
    ~EarlyScope();
    const EarlyParams &params() const { return p_; }
    void done(bool launched);
private:
    EarlyParams p_;
    void *rec_;    // the (device, stream) record of b200dct.cu
    Span w_[2];
    bool feeds_, finished_;
};
void note_launch(int launches, const char *path); // b200dct_last_launch_count / b200dct_last_path of this thread
bool use_factored_inverse_u8(const b200dct_plan *pl);
bool pdl_enabled(cudaStream_t s);
void forget_stream(cudaStream_t s); // a non-TMA kernel of the library was launched on s (see early loads, b200dct.cu)
}
