// field.hpp -- host array with ghost cells, same members and semantics as the reference's Field2D
// (src/solver/field.hpp:9-73): 64-byte aligned, zero-filled, move-only, throws std::bad_alloc.
// The B200 path allocates it as PINNED host memory when a CUDA runtime is reachable (cfd_b200_host_alloc), so the
// State <-> device copies of the drop-in run at full PCIe speed; it falls back to aligned_alloc otherwise.
#pragma once

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

extern "C" void *cfd_b200_host_alloc(size_t bytes); // pinned, 64-byte aligned; NULL on failure
extern "C" void cfd_b200_host_free(void *p);
namespace b200 { void forget_host_array(const void *data); } // a host array goes away: no Session may keep mirroring it

template <typename T> struct Field2D {
    int nx = 0, ny = 0, pitch = 0, ngx = 0, ngy = 0;
    T *data = nullptr;

    Field2D() = default;
    ~Field2D() { free(); }
    Field2D(const Field2D &) = delete;
    Field2D &operator=(const Field2D &) = delete;
    Field2D(Field2D &&o) noexcept { swap(o); }
    Field2D &operator=(Field2D &&o) noexcept {
        if (this != &o) {
            free();
            swap(o);
        }
        return *this;
    }

    void allocate(int NX, int NY, int pitch_, int ngx_, int ngy_) {
        free();
        nx = NX, ny = NY, pitch = pitch_, ngx = ngx_, ngy = ngy_;
        const size_t bytes = count() * sizeof(T);
        data = static_cast<T *>(cfd_b200_host_alloc(bytes));
        if (!data) {
            nx = ny = pitch = ngx = ngy = 0;
            throw std::bad_alloc();
        }
        std::memset(data, 0, bytes);
    }

    void free() {
        if (data) b200::forget_host_array(data), cfd_b200_host_free(data);
        data = nullptr;
        nx = ny = pitch = ngx = ngy = 0;
    }

    size_t count() const { return static_cast<size_t>(pitch) * (ny + 2 * ngy); }
    T &at_raw(int ii, int jj) { return data[jj * pitch + ii]; }
    const T &at_raw(int ii, int jj) const { return data[jj * pitch + ii]; }

  private:
    void swap(Field2D &o) {
        std::swap(nx, o.nx), std::swap(ny, o.ny), std::swap(pitch, o.pitch);
        std::swap(ngx, o.ngx), std::swap(ngy, o.ngy), std::swap(data, o.data);
    }
};
