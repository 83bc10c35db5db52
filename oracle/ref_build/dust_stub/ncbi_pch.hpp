/* stand-in for the toolkit's precompiled header (oracle/Makefile, target dust) */
