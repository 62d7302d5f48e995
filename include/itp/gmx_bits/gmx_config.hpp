// gmx_main(func): defines main() that hands control to `func` through gmx_run_cmain, as the reference macro does
// (reference include/itp/gmx_bits/gmx_config.hpp:3-9), so `gmx_main(temp) { ... }` programs compile unchanged.
#pragma once
#define gmx_main(func)                                                                                             \
    int func(int argc, char** argv);                                                                               \
    int main(int argc, char** argv) { return gmx_run_cmain(argc, argv, &func); }                                   \
    int func(int argc, char** argv)
