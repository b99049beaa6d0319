// helmnet_emu.cpp -- TEST INFRASTRUCTURE ONLY: the product sources compiled against the fiber emulator.
#define HN_EMU 1
#include "../../helmnet_b200/csrc/helmnet_sm100.cu"
