#pragma once
#include "Bmp.h"
