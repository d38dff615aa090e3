#pragma once
#include "global_managers.hpp"
