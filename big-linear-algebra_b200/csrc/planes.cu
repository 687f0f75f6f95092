// planes.cu -- see planes.h
#include "planes.h"

#include <cstring>

namespace bla {

PlaneSet::PlaneSet(CallScope& sc, Matrix* planes, int count, bool read) : sc_(sc) {
    ptrs_.reserve(count);
    for (int i = 0; i < count; ++i) ptrs_.push_back(planes[i].data);
    plane_ = count ? (size_t)planes[0].rows * planes[0].cols : 0;
    setup(read);
}

PlaneSet::PlaneSet(CallScope& sc, Matrix** table, int F, int C, bool read) : sc_(sc) {
    ptrs_.reserve((size_t)F * C);
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < C; ++c) ptrs_.push_back(table[f][c].data);
    plane_ = (F && C) ? (size_t)table[0][0].rows * table[0][0].cols : 0;
    setup(read);
}

void PlaneSet::setup(bool read) {
    const size_t count = ptrs_.size();
    if (!count || !plane_) return;
    const size_t bytes = count * plane_ * sizeof(float);
    MemKind first = classify(ptrs_.front()), last = classify(ptrs_.back());
    host_ = (first == kHost || first == kPinned) && (last == kHost || last == kPinned);
    sc_.note_kind(host_ ? kHost : first);
    if (!host_) sc_.note_kind(last);
    bool contiguous = true;
    for (size_t i = 1; i < count && contiguous; ++i) contiguous = ptrs_[i] == ptrs_[0] + i * plane_;
    if (!host_ && contiguous) {
        in_place_ = true;
        dev_ = ptrs_[0];
        return;
    }
    dev_ = (float*)sc_.scratch_bytes(bytes);
    if (host_) {
        pinned_ = (float*)pool_alloc(kPinned, bytes);
        if (read) {
            for (size_t i = 0; i < count; ++i) memcpy(pinned_ + i * plane_, ptrs_[i], plane_ * sizeof(float));
            BLA_CUDA(cudaMemcpyAsync(dev_, pinned_, bytes, cudaMemcpyHostToDevice, sc_.stream()));
            rt().h2d_bytes += bytes;
        }
    } else if (read) {
        for (size_t i = 0; i < count; ++i)
            BLA_CUDA(cudaMemcpyAsync(dev_ + i * plane_, ptrs_[i], plane_ * sizeof(float), cudaMemcpyDefault, sc_.stream()));
    }
}

void PlaneSet::write_back() {
    const size_t count = ptrs_.size();
    if (in_place_ || !count || !plane_) return;
    const size_t bytes = count * plane_ * sizeof(float);
    if (host_) {
        BLA_CUDA(cudaMemcpyAsync(pinned_, dev_, bytes, cudaMemcpyDeviceToHost, sc_.stream()));
        rt().d2h_bytes += bytes;
        sc_.need_sync();
        pending_scatter_ = true;
    } else {
        for (size_t i = 0; i < count; ++i)
            BLA_CUDA(cudaMemcpyAsync(ptrs_[i], dev_ + i * plane_, plane_ * sizeof(float), cudaMemcpyDefault, sc_.stream()));
    }
}

void PlaneSet::finish_host_scatter() {
    if (!pending_scatter_) return;
    pending_scatter_ = false;
    sc_.finish();   // synchronises the stream (idempotent)
    for (size_t i = 0; i < ptrs_.size(); ++i) memcpy(ptrs_[i], pinned_ + i * plane_, plane_ * sizeof(float));
}

PlaneSet::~PlaneSet() {
    finish_host_scatter();
    if (pinned_) {
        if (!pending_scatter_) {
            // a read-only pinned block may still be the source of an in-flight H2D copy
            BLA_CUDA(cudaStreamSynchronize(sc_.stream()));
        }
        pool_free(pinned_);
    }
}

}  // namespace bla
