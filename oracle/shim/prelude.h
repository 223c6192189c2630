// Force-included before the reference headers: std headers that libc++ (the reference's native
// toolchain, clang++) provides transitively and libstdc++ does not.  Test infrastructure only.
#pragma once
#include <algorithm>
#include <bit>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <tuple>
#include <typeindex>
#include <utility>
#include <variant>
