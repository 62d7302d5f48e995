// calc_density_distribution_in_pore (reference src/mofs/calc_density_distribution_in_pore.cpp), ported to the device.
//
// Same command line and output file.  The frame loop (:45-61: centres, z slab test, cylinder test, one histogram update per
// molecule) becomes one b2aAccumulate() of a region accumulator per resident batch; the normalisation (:63-70) is the
// reference's arithmetic on the returned counts.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);

    int  nbin = 100;
    real lowPos = 7.761, upPos = 12.239, centerX = 2.829, centerY = 2.056, Rmof = 0.927;
    hd.pa = {
        { "-nbin", FALSE, etINT, { &nbin }, "number of histogram bins" },
        { "-upPos", FALSE, etREAL, { &upPos }, "upper bound of the pore region [nm]" },
        { "-lowPos", FALSE, etREAL, { &lowPos }, "lower bound of the pore region [nm]" },
        { "-centerX", FALSE, etREAL, { &centerX }, "pore axis, x coordinate [nm]" },
        { "-centerY", FALSE, etREAL, { &centerY }, "pore axis, y coordinate [nm]" },
        { "-Rmof", FALSE, etREAL, { &Rmof }, "pore radius [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "numberDensity", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const double Rsquare = double(Rmof) * Rmof;          // :38
    const double dbin    = (upPos - lowPos) / nbin;      // :41, float arithmetic widened

    b2a_region_spec spec{};
    spec.grp = 0; spec.com = B2A_COM_MASS;
    spec.nfilter = 1; spec.fdim[0] = ZZ; spec.flow[0] = lowPos; spec.fup[0] = upPos; spec.fincl[0] = 0;   // lowPos <= z < upPos
    spec.cyl = 1; spec.cx = centerX; spec.cy = centerY; spec.R2 = Rsquare;
    spec.nbdim = 1; spec.bdim[0] = ZZ; spec.blow[0] = lowPos; spec.dbin[0] = dbin; spec.nb[0] = nbin;
    const int acc = hd.b2aRegion(spec);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    itp::vecd density(nbin);
    for (int i = 0; i < nbin; ++i) density[i] = (double)counts[i];
    const double dvol = dbin * itp::pi * Rmof * Rmof;    // :63
    density /= dvol;

    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    for (int i = 0; i != nbin; ++i) fprintf(ofile, "%8.4e %8.4e\n", lowPos + dbin / 2 + i * dbin, density[i] / hd.nframe);
    fclose(ofile);
    return 0;
}
