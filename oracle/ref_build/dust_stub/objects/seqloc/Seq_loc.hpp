/* stand-in: a Seq-loc is an id plus [from, to] here */
#ifndef DUST_STUB_SEQ_LOC_HPP
#define DUST_STUB_SEQ_LOC_HPP
#include <corelib/ncbiobj.hpp>
BEGIN_NCBI_SCOPE
BEGIN_SCOPE(objects)
class CSeq_id {};
class CSeq_loc {
public:
    CSeq_loc(CSeq_id &, unsigned from, unsigned to) : from_(from), to_(to) {}
    unsigned from_, to_;
};
END_SCOPE(objects)
END_NCBI_SCOPE
#endif
