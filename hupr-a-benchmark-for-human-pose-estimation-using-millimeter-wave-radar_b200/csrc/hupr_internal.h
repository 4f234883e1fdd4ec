// Internal include: pulls the public C ABI declarations into the kernel translation units.
#pragma once
#include "hupr_b200.h"
