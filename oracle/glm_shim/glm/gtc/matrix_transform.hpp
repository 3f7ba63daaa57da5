// Stand-in header: everything lives in glm/glm.hpp (see that file).
#pragma once
#include "../glm.hpp"
