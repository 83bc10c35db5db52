/* stand-in: scope macros only */
#ifndef DUST_STUB_NCBISTR_HPP
#define DUST_STUB_NCBISTR_HPP
#define BEGIN_NCBI_SCOPE namespace ncbi {
#define END_NCBI_SCOPE }
#define BEGIN_SCOPE(x) namespace x {
#define END_SCOPE(x) }
#define NCBI_XALGODUSTMASK_EXPORT
#include <algorithm>
namespace ncbi { using std::max; using std::min; }      /* the toolkit's ncbistd brings these into its namespace */
#endif
