/* case shim: the reference includes "Core.h", the file on disk is core.h */
#pragma once
#include "core.h"
