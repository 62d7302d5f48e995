// calc_radius_distribution_in_pore (reference src/mofs/calc_radius_distribution_in_pore.cpp), ported to the device.
//
// Same command line and output file.  The frame loop (:43-54: centres, z slab test, cylinder test, histogram of the distance from
// the pore axis) becomes one b2aAccumulate() of a region accumulator with a radial binned dimension per resident batch; the shell
// volumes and the normalisation (:55-70) are the reference's arithmetic.
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
    hd.fnm = { { efXVG, "-o", "rdf_in_pore", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const double Rsquare = Rmof * Rmof;     // float product widened, :38
    const double dbin    = Rmof / nbin;     // :40

    b2a_region_spec spec{};
    spec.grp = 0; spec.com = B2A_COM_MASS;
    spec.nfilter = 1; spec.fdim[0] = ZZ; spec.flow[0] = lowPos; spec.fup[0] = upPos; spec.fincl[0] = 0;
    spec.cyl = 1; spec.cx = centerX; spec.cy = centerY; spec.R2 = Rsquare;
    spec.nbdim = 1; spec.bdim[0] = 3; spec.blow[0] = 0; spec.dbin[0] = dbin; spec.nb[0] = nbin;   // meZ = sqrt(rsquare) / dbin
    const int acc = hd.b2aRegion(spec);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    itp::vecd rdf(nbin), volumn(nbin);
    for (int i = 0; i < nbin; ++i) rdf[i] = (double)counts[i];
    volumn.fill(0);
    for (size_t i = 0; i != (size_t)nbin; ++i) volumn[i] = dbin * (i + 1) * dbin * (i + 1);
    for (size_t i = nbin - 1; i != 0; --i) volumn[i] = volumn[i] - volumn[i - 1];
    volumn *= (itp::pi * (upPos - lowPos));
    rdf /= volumn;

    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    for (size_t i = 0; i != (size_t)nbin; ++i) fprintf(ofile, "%8.4e %8.4e\n", dbin / 2 + i * dbin, rdf[i] / hd.nframe);
    fclose(ofile);
    return 0;
}
