// TEST INFRASTRUCTURE: boost::alignment::aligned_allocator<T, A> as src/signal.h uses it (std::vector storage).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <new>
namespace boost { namespace alignment {
template <class T, std::size_t A> struct aligned_allocator {
    typedef T value_type;
    aligned_allocator() {}
    template <class U> aligned_allocator(const aligned_allocator<U, A> &) {}
    template <class U> struct rebind { typedef aligned_allocator<U, A> other; };
    T *allocate(std::size_t n) {
        void *p = std::aligned_alloc(A, ((n * sizeof(T) + A - 1) / A) * A);
        if (!p) throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, std::size_t) { std::free(p); }
    template <class U> bool operator==(const aligned_allocator<U, A> &) const { return true; }
    template <class U> bool operator!=(const aligned_allocator<U, A> &) const { return false; }
};
} }
