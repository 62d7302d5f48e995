// The part of the GROMACS 2020 C API the handles are written against (reference include/itp/gmx:9-19 pulls in the same
// eleven headers).  With a real installation these resolve to libgromacs' headers, in this image to gmxstub/include.
#pragma once

// command line, file-name options, program entry (parse_common_args, ftp2fn / opt2fn, gmx_run_cmain)
#include <gromacs/commandline/cmdlineinit.h>
#include <gromacs/commandline/pargs.h>
// topology and index groups (read_tpx_top, get_index, t_topology, t_inputrec)
#include <gromacs/fileio/tpxio.h>
#include <gromacs/mdtypes/inputrec.h>
#include <gromacs/topology/index.h>
#include <gromacs/topology/topology.h>
// trajectory frames (read_first_frame / read_next_frame, t_trxframe, TRX_READ_*)
#include <gromacs/fileio/trxio.h>
#include <gromacs/trajectory/trajectoryframe.h>
// utilities the programs use directly (gmx_fatal / FARGS, snew, XX / YY / ZZ, real, rvec)
#include <gromacs/math/utilities.h>
#include <gromacs/utility/fatalerror.h>
#include <gromacs/utility/smalloc.h>
