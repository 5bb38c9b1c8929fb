#pragma once
#include "tree.h"
