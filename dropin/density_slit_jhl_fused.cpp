// density_slit_jhl (reference src/general/density_slit_jhl.cpp), ported to the fused device accumulator.
//
// Same command line and output file.  The frame loop (:54-67: loadPositionCenter, two inclusive region tests, one histogram
// update per molecule) becomes one b2aAccumulate() per resident batch (K2 with upper_inclusive = 1 and the second filter
// dimension); the normalisation by frames x slab volume (:69-74) is the reference's arithmetic on the returned counts.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);
    hd.ngrps = 1;

    int  nbin1 = 100, dim1 = 0, dim2 = 2;
    real lowPos1 = 0, upPos1 = 30, upPos2 = 0, lowPos2 = 10, yy = 0;
    hd.pa = {
        { "-nbin1", FALSE, etINT, { &nbin1 }, "number of histogram bins" },
        { "-dim1", FALSE, etINT, { &dim1 }, "axis of the slab: 0 = x, 1 = y, 2 = z" },
        { "-dim2", FALSE, etINT, { &dim2 }, "axis of the slab: 0 = x, 1 = y, 2 = z" },
        { "-up1", FALSE, etREAL, { &upPos1 }, "upper bound of the selection slab [nm]" },
        { "-low1", FALSE, etREAL, { &lowPos1 }, "lower bound of the selection slab [nm]" },
        { "-up2", FALSE, etREAL, { &upPos2 }, "upper bound of the selection slab [nm]" },
        { "-low2", FALSE, etREAL, { &lowPos2 }, "lower bound of the selection slab [nm]" },
        { "-yy", FALSE, etREAL, { &yy }, "box length along y [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "density_slit", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const double dbin = (upPos1 - lowPos1) / nbin1;   // float arithmetic widened, :52

    b2a_density_spec spec{};
    spec.grp = 0; spec.com = B2A_COM_MASS; spec.dim = dim1; spec.low = lowPos1; spec.up = upPos1; spec.dbin = dbin; spec.nbin = nbin1;
    spec.upper_inclusive = 1;   // lowPos <= p <= upPos on both dimensions (:60-61)
    spec.dim2 = dim2; spec.low2 = lowPos2; spec.up2 = upPos2;
    const int acc = hd.b2aDensity(spec);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    auto         ofile   = hd.openWrite(hd.get_opt2fn("-o"));
    const double dVolume = (upPos2 - lowPos2) * dbin * yy;
    for (int i = 0; i < nbin1; i++)
    {
        double d = (double)counts[i];
        d /= (hd.nframe * dVolume);   // #/nm^3, :72
        fmt::print(ofile, "{:15.8f} {:15.8f}\n", dbin / 2 + dbin * i, d);
    }
    fclose(ofile);
    return 0;
}
