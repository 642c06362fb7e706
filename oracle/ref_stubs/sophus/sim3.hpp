#include "../matcher_world.h"
