#pragma once
#include "mathlib/vector.h"
