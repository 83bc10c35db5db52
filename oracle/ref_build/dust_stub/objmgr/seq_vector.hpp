/* stand-in for CSeqVector / CSeqVector_CI: a string of IUPACNA letters with the iterator calls symdust.cpp makes
 * (construction at a position, *, ++, SetPos, GetPos) */
#ifndef DUST_STUB_SEQ_VECTOR_HPP
#define DUST_STUB_SEQ_VECTOR_HPP
#include <string>
#include <corelib/ncbiobj.hpp>
BEGIN_NCBI_SCOPE
BEGIN_SCOPE(objects)
class CSeqVector;
class CSeqVector_CI {
public:
    CSeqVector_CI(const CSeqVector &v, unsigned pos);
    char operator*() const;
    CSeqVector_CI &operator++() { ++pos_; return *this; }
    void SetPos(unsigned p) { pos_ = p; }
    unsigned GetPos() const { return pos_; }
private:
    const CSeqVector *v_;
    unsigned pos_;
};
class CSeqVector {
public:
    typedef unsigned size_type;
    typedef CSeqVector_CI const_iterator;
    explicit CSeqVector(const std::string &s) : s_(s) {}
    size_type size() const { return (size_type)s_.size(); }
    bool empty() const { return s_.empty(); }
    char at(unsigned p) const { return p < s_.size() ? s_[p] : 'N'; }
private:
    std::string s_;
};
inline CSeqVector_CI::CSeqVector_CI(const CSeqVector &v, unsigned pos) : v_(&v), pos_(pos) {}
inline char CSeqVector_CI::operator*() const { return v_->at(pos_); }
END_SCOPE(objects)
END_NCBI_SCOPE
#endif
