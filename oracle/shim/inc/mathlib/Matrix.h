#pragma once
#include "mathlib/matrix.h"
