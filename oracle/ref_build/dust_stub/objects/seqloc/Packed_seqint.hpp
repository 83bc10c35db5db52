#ifndef DUST_STUB_PACKED_SEQINT_HPP
#define DUST_STUB_PACKED_SEQINT_HPP
#include <vector>
#include <utility>
#include <objects/seqloc/Seq_loc.hpp>
BEGIN_NCBI_SCOPE
BEGIN_SCOPE(objects)
class CPacked_seqint {
public:
    void AddInterval(CSeq_id &, unsigned from, unsigned to) { ivs.push_back(std::make_pair(from, to)); }
    std::vector<std::pair<unsigned, unsigned> > ivs;
};
END_SCOPE(objects)
END_NCBI_SCOPE
#endif
