// planes.h -- the reference passes images as ARRAYS of separately allocated channel planes
// (`Matrix* X`, model/cifar_unet.c:317-322) and kernels as `Matrix**` tables of k x k planes.
// On the device every tensor is one contiguous [C][H][W] (or [F][C][k][k]) block; these helpers
// move between the two views.  Host planes are packed into one pinned staging block and cross
// PCIe in a single copy; device planes that already sit back to back are used in place.
#pragma once
#include "../../include/lib/matrix.h"
#include "runtime.h"

namespace bla {

class PlaneSet {
public:
    // planes[0..count) all rows x cols.  `read`: contents are needed on the device.
    PlaneSet(CallScope& sc, Matrix* planes, int count, bool read);
    // kernels[f][c], F x C planes of k x k
    PlaneSet(CallScope& sc, Matrix** table, int F, int C, bool read);
    float* dev() const { return dev_; }
    size_t plane_elems() const { return plane_; }
    // queue the copy of the device block back into the caller's planes (call before sc.finish())
    void write_back();
    // must be called after the stream has been synchronised when write_back() was host-bound
    void finish_host_scatter();
    ~PlaneSet();

private:
    CallScope& sc_;
    std::vector<float*> ptrs_;
    size_t plane_ = 0;
    float* dev_ = nullptr;
    float* pinned_ = nullptr;
    bool in_place_ = false;
    bool host_ = false;
    bool pending_scatter_ = false;
    void setup(bool read);
};

}  // namespace bla
