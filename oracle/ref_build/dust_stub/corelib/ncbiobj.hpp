/* stand-in: CRef / CConstRef as plain owning pointers (enough for symdust.cpp's two convenience wrappers) */
#ifndef DUST_STUB_NCBIOBJ_HPP
#define DUST_STUB_NCBIOBJ_HPP
#include <memory>
#include "ncbistr.hpp"
BEGIN_NCBI_SCOPE
template <class T> class CRef {
public:
    CRef() {}
    explicit CRef(T *p) : p_(p) {}
    T *operator->() const { return p_.get(); }
    T &operator*() const { return *p_; }
private:
    std::shared_ptr<T> p_;
};
template <class T> class CConstRef {
public:
    CConstRef() {}
    explicit CConstRef(const T *p) : p_(p) {}
    const T *operator->() const { return p_.get(); }
private:
    std::shared_ptr<const T> p_;
};
END_NCBI_SCOPE
#endif
