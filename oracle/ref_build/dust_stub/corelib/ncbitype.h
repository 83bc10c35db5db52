/* stand-in: the fixed-width names symdust uses (the toolkit's corelib/ncbitype.h needs the configured build tree) */
#ifndef DUST_STUB_NCBITYPE_H
#define DUST_STUB_NCBITYPE_H
#include <stdint.h>
typedef uint8_t Uint1;
typedef uint32_t Uint4;
typedef int32_t Int4;
#endif
