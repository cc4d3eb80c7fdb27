// stand-in: src/LightGlue.cc includes it next to NvInfer.h.  TEST INFRASTRUCTURE.
#pragma once
#include "NvInfer.h"
