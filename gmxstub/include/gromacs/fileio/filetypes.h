/* Stand-in for <gromacs/fileio/filetypes.h> (libgromacs 2020.1 is not installed in this image).
 * Every stand-in header forwards to gmxstub_api.h, which declares the subset of the
 * GROMACS C API that include/itp/gmx and the analysis programs use. */
#pragma once
#include "gromacs/gmxstub_api.h"
