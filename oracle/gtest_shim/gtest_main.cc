// TEST INFRASTRUCTURE — main() for binaries built with the gtest shim.
#include "gtest/gtest.h"
int main() { return ::testing::RunAllTests(); }
