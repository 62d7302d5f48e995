// Exercises the GmxHandleFull surface (init, frame stepping, initPos/initPosc, loadPosition, loadPositionCenter,
// loadVelocityCenter with -vel 1) and
// dumps every value, so that the same source compiled against the reference header and against the drop-in header can
// be compared byte for byte (tests/test_gpu_dropin.py).  Not a reference program: the reference's GmxHandleFull users
// (mofs/calc_qmsd_pore3*.cpp, calc_ecacf-in-pore.cpp) compute MSD / current ACFs, which are out of scope.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandleFull hd(argc, argv);
    hd.ngrps = 2;
    int center = 0, vel = 0;
    hd.pa = { { "-center", FALSE, etINT, { &center }, "0 mass, 2 charge" }, { "-vel", FALSE, etINT, { &vel }, "also dump velocity centres" } };
    hd.fnm = { { efDAT, "-o", "full_dump", ffWRITE } };
    hd.init();
    if (vel) hd.flags = TRX_READ_X | TRX_READ_V;
    hd.readFirstFrame();
    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    fmt::print(ofile, "nmol {} {} first napm {} {}\n", hd.nmol[0], hd.nmol[1], hd.napm[0][0], hd.napm[1][0]);
    itp::boxd pos;
    itp::matd posc, velc;
    do
    {
        for (int g = 0; g < 2; ++g)
        {
            hd.initPos(pos, g);
            for (Eigen::Index i = 0; i < pos.size(); ++i) pos.data()[i].setZero();
            hd.initPosc(posc, g);
            hd.loadPosition(pos, g);
            hd.loadPositionCenter(posc, g, center);
            fmt::print(ofile, "frame {} time {:.3f} dt {:.3f} grp {}\n", hd.nframe, hd.time, hd.dt, g);
            if (vel)
            {
                hd.initPosc(velc, g);
                hd.loadVelocityCenter(velc, g, center);
                for (int i = 0; i < hd.nmol[g]; ++i) fmt::print(ofile, "v {:.17g} {:.17g} {:.17g}\n", velc(i, 0), velc(i, 1), velc(i, 2));
            }
            for (int i = 0; i < hd.nmol[g]; ++i)
            {
                fmt::print(ofile, "{:.17g} {:.17g} {:.17g}", posc(i, 0), posc(i, 1), posc(i, 2));
                for (int j = 0; j < hd.napm[g][i]; ++j) fmt::print(ofile, " {:.17g} {:.17g} {:.17g}", pos(i, j)[0], pos(i, j)[1], pos(i, j)[2]);
                fmt::print(ofile, "\n");
            }
        }
    } while (hd.readNextFrame());
    fclose(ofile);
    return 0;
}
