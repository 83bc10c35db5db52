/* Stand-in for the configure-generated ncbiconf_unix.h of the NCBI C++ Toolkit.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): lets the reference's own C
 * sources under /root/reference compile in place with plain gcc, without
 * running the reference's build system. Written for x86-64 Linux / gcc. */
#ifndef GBLASTN_B200_ORACLE_NCBICONF_UNIX_H
#define GBLASTN_B200_ORACLE_NCBICONF_UNIX_H
#define NCBI_OS_UNIX 1
#define NCBI_OS_LINUX 1
#define NCBI_OS "linux-gnu"
#define NCBI_COMPILER_GCC 1
#define NCBI_COMPILER_VERSION 1330
#define HOST "x86_64-unknown-linux-gnu"
#define HOST_CPU "x86_64"
#define HOST_VENDOR "unknown"
#define HOST_OS "linux-gnu"
#define SIZEOF_CHAR 1
#define SIZEOF_SHORT 2
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define SIZEOF___INT64 0
#define SIZEOF_FLOAT 4
#define SIZEOF_DOUBLE 8
#define SIZEOF_LONG_DOUBLE 16
#define SIZEOF_SIZE_T 8
#define SIZEOF_VOIDP 8
#define HAVE_STDINT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_LIMITS_H 1
#define HAVE_STRING_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STRDUP 1
#define HAVE_STRNDUP 1
#define HAVE_STRCASECMP 1
#define HAVE_ERF 1
#define HAVE_ATTRIBUTE_DESTRUCTOR 1
#define STDC_HEADERS 1
#define NCBI_THREADS 1
#define NCBI_POSIX_THREADS 1
#endif
