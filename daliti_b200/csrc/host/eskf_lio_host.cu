#include "eskf_lio_host.hpp"
