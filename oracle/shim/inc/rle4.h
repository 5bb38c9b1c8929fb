#pragma once
#include "Rle4.h"
