// emu_main.cpp — TEST INFRASTRUCTURE: builds the product's kernel sources against the SIMT emulator
// (tests/emu/simt_emu.h) into tests/emu/libconsent_emu.so.  Loaded only by `-m "not gpu"` tests.
// (built with -DCG_EMU -DCG_EMU_IMPL -include simt_emu.h)
#include "simt_emu.h"
#include "../../consent_b200/csrc/consent_b200.cu"
