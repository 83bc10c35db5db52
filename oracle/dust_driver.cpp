// dust_driver.cpp — C entry point around the reference's own symmetric DUST (TEST INFRASTRUCTURE).
//
// c++/src/algo/dustmask/symdust.cpp is compiled where it lies in /root/reference (oracle/Makefile, target `dust`)
// against the stand-in headers of oracle/ref_build/dust_stub/: the algorithm is the reference's, the object-manager
// types it is written against (CSeqVector, CSeq_loc, CRef) are replaced by minimal ones because the real ones need
// the toolkit's configured build tree.  Used by tests/test_dust.py to pin gblastn_b200's bn_dust_mask.
#include <cstdlib>
#include <memory>
#include <string>
#include <algo/dustmask/symdust.hpp>

extern "C" int ref_dust(const char *iupac, int len, int level, int window, int linker, int **out, int *n)
{
    ncbi::objects::CSeqVector seq(std::string(iupac, (size_t)len));
    ncbi::CSymDustMasker masker((Uint4)level, (unsigned)window, (unsigned)linker);
    std::auto_ptr<ncbi::CSymDustMasker::TMaskList> res = masker(seq);
    *n = (int)res->size();
    *out = (int *)malloc(sizeof(int) * 2 * (size_t)(*n ? *n : 1));
    for (int i = 0; i < *n; i++) { (*out)[2 * i] = (int)(*res)[i].first; (*out)[2 * i + 1] = (int)(*res)[i].second; }
    return 0;
}
extern "C" void ref_dust_free(int *p) { free(p); }
