// Empty stand-in for boost/pool/object_pool.hpp (include-only use in the reference).
#pragma once
