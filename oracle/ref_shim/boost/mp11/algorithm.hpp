#pragma once
#include <boost/mp11.hpp>
