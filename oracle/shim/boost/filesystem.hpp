#pragma once
#include "boost/filesystem/operations.hpp"
