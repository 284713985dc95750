// forces_prelude.h — TEST INFRASTRUCTURE ONLY; force-included (-include) before /root/reference/src/Forces.cpp, UtilEOL.cpp and
// conversions.cpp when oracle/Makefile compiles them UNMODIFIED into oracle/_ref/libforces_ref.so.
//
// Why: src/Cloth.h:28 and :37 read `extern struct Material {` / `extern struct Remeshing {` — a storage class on a type
// definition, which MSVC accepts and g++ rejects whatever the flags (error: "a storage class can only be specified for objects
// and functions").  Forces.h:9 includes Cloth.h only for `struct Material`.  The build therefore pre-defines Cloth.h's include
// guard (-D__Cloth__, Cloth.h:2-3) so that the reference's own `#ifndef __Cloth__` skips the file, and this prelude declares the
// one thing Forces.cpp needs from it, field for field as Cloth.h:28-35 has it.  No reference source is edited or copied.
#pragma once
struct Material {
	double density; // area density
	double e;
	double nu;
	double beta;
	double dampingA;
	double dampingB;
};
