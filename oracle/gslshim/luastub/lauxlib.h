#include "lua.h"
