// TEST INFRASTRUCTURE. Minimal stand-in for boost::circular_buffer (Boost is not in this image),
// written from Boost's documented semantics, so that /root/reference/src/utils.h (MovingAverage,
// DCBlocker) compiles UNMODIFIED into oracle/_ref/libphantom_ref.so. Only the members utils.h uses:
//   circular_buffer(capacity, value)  -> a FULL buffer holding `capacity` copies of value
//   push_front(v)                     -> on a full buffer overwrites the back element
//   back(), operator[](i) (0 = front/newest after push_front), begin()/end(), size()
#pragma once
#include <cstddef>
#include <iterator>
#include <vector>

namespace boost {
template <typename T> class circular_buffer {
  public:
    circular_buffer() : head_(0), size_(0) {}
    circular_buffer(std::size_t capacity, const T &value) : data_(capacity, value), head_(0), size_(capacity) {}
    std::size_t size() const { return size_; }
    std::size_t capacity() const { return data_.size(); }
    T &operator[](std::size_t i) { return data_[(head_ + i) % data_.size()]; }
    const T &operator[](std::size_t i) const { return data_[(head_ + i) % data_.size()]; }
    T &front() { return (*this)[0]; }
    T &back() { return (*this)[size_ - 1]; }
    void push_front(const T &v) {
        if (data_.empty()) return;
        head_ = (head_ + data_.size() - 1) % data_.size();
        data_[head_] = v;
        if (size_ < data_.size()) size_++;
    }
    class iterator {
      public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = T;
        using difference_type = std::ptrdiff_t;
        using pointer = T *;
        using reference = T &;
        iterator(circular_buffer *b, std::size_t i) : b_(b), i_(i) {}
        reference operator*() const { return (*b_)[i_]; }
        iterator &operator++() { ++i_; return *this; }
        iterator operator++(int) { iterator t = *this; ++i_; return t; }
        bool operator==(const iterator &o) const { return i_ == o.i_; }
        bool operator!=(const iterator &o) const { return i_ != o.i_; }
      private:
        circular_buffer *b_;
        std::size_t i_;
    };
    iterator begin() { return iterator(this, 0); }
    iterator end() { return iterator(this, size_); }
  private:
    std::vector<T> data_;
    std::size_t head_, size_;
};
} // namespace boost
