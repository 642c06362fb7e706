#include "parallel_for.h"
