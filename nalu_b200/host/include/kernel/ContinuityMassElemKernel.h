#pragma once
#include "MassElemKernels.h"
