#pragma once
#include "ScalarAdvDiffElemKernel.h"
