// TEST INFRASTRUCTURE ONLY: stand-in for <Windows.h> so that the reference decoder (decoder/YAIK_API.cpp) compiles on
// Linux in its YAIK_DEVEL configuration (its release configuration does not compile: DebugRGBAsPng is undeclared and
// `checkAllocationEnd();` expands to `();`).  See oracle/Makefile, target dec.
#pragma once
#define _CrtCheckMemory() ((void)0)
