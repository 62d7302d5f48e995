// calc_areaDensityDistributionPore (reference src/mofs/calc_areaDensityDistributionPore.cpp), ported to the device.
//
// Same command line and output file.  The frame loop (:54-68: centres, three half-open region tests, one update of a 2-D
// histogram per molecule) becomes one b2aAccumulate() of a region accumulator with two binned dimensions per resident batch.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);

    int  dim1 = 0, dim2 = 1, dim3 = 2;
    real lowPos1 = 1.66, upPos1 = 3.996, lowPos2 = 1.133, upPos2 = 3.073, lowPos3 = 7.761, upPos3 = 12.239, dbin = 0.002;
    hd.pa = {
        { "-dim1", FALSE, etINT, { &dim1 }, "axis of filter 1 (0 = x, 1 = y, 2 = z)" },
        { "-up1", FALSE, etREAL, { &upPos1 }, "upper bound of filter 1 [nm]" },
        { "-low1", FALSE, etREAL, { &lowPos1 }, "lower bound of filter 1 [nm]" },
        { "-dim2", FALSE, etINT, { &dim2 }, "axis of filter 2 (0 = x, 1 = y, 2 = z)" },
        { "-up2", FALSE, etREAL, { &upPos2 }, "upper bound of filter 2 [nm]" },
        { "-low2", FALSE, etREAL, { &lowPos2 }, "lower bound of filter 2 [nm]" },
        { "-dim3", FALSE, etINT, { &dim3 }, "axis of filter 3 (0 = x, 1 = y, 2 = z)" },
        { "-up3", FALSE, etREAL, { &upPos3 }, "upper bound of filter 3 [nm]" },
        { "-low3", FALSE, etREAL, { &lowPos3 }, "lower bound of filter 3 [nm]" },
        { "-dbin", FALSE, etREAL, { &dbin }, "bin width [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "areaDensity", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const int nbin1 = (upPos1 - lowPos1) / dbin;   // :47-48, float arithmetic
    const int nbin2 = (upPos2 - lowPos2) / dbin;

    b2a_region_spec spec{};
    spec.grp = 0; spec.com = B2A_COM_MASS;
    spec.nfilter = 3;
    spec.fdim[0] = dim1; spec.flow[0] = lowPos1; spec.fup[0] = upPos1;
    spec.fdim[1] = dim2; spec.flow[1] = lowPos2; spec.fup[1] = upPos2;
    spec.fdim[2] = dim3; spec.flow[2] = lowPos3; spec.fup[2] = upPos3;
    spec.nbdim = 2;
    spec.bdim[0] = dim1; spec.blow[0] = lowPos1; spec.dbin[0] = dbin; spec.nb[0] = nbin1;   // dens(me1, me2): me1 + me2 * nbin1
    spec.bdim[1] = dim2; spec.blow[1] = lowPos2; spec.dbin[1] = dbin; spec.nb[1] = nbin2;
    const int acc = hd.b2aRegion(spec);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    itp::matd dens(nbin1, nbin2);
    for (int j = 0; j != nbin2; ++j)
        for (int i = 0; i != nbin1; ++i) dens(i, j) = (double)counts[(size_t)i + (size_t)j * nbin1];
    dens /= (dbin * dbin * (upPos3 - lowPos3) * hd.nframe);   // :70

    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    fmt::print(ofile, "# #/nm^3\n");
    for (int i = 0; i != nbin1; ++i)
    {
        for (int j = 0; j != nbin2; ++j) fmt::print(ofile, "{:8.4f} ", dens(i, j));
        fmt::print(ofile, "\n");
    }
    fclose(ofile);
    return 0;
}
