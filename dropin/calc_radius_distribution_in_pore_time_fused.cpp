// calc_radius_distribution_in_pore_time (reference src/mofs/calc_radius_distribution_in_pore_time.cpp), ported to the device.
//
// Same command line and output file: for every frame, the histogram of the distance from the pore axis of the molecules inside a
// box, divided by the shell volumes.  The per-frame loop (:80-93) becomes one b2aAccumulate() of a region accumulator (three
// half-open filters, radial binned dimension, per-frame counters) per resident batch.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);

    real x1 = 0, x2 = 0, y1 = 0, y2 = 0, z1 = 0, z2 = 0, centerX = 2.829, centerY = 2.056, dr = 0.01, Cr = 0, CR = 0.5;
    int  dim = 2;
    hd.pa = {
        { "-x1", FALSE, etREAL, { &x1 }, "pore box: lower x bound [nm]" },
        { "-x2", FALSE, etREAL, { &x2 }, "pore box: upper x bound [nm]" },
        { "-y1", FALSE, etREAL, { &y1 }, "pore box: lower y bound [nm]" },
        { "-y2", FALSE, etREAL, { &y2 }, "pore box: upper y bound [nm]" },
        { "-z1", FALSE, etREAL, { &z1 }, "pore box: lower z bound [nm]" },
        { "-z2", FALSE, etREAL, { &z2 }, "pore box: upper z bound [nm]" },
        { "-centerX", FALSE, etREAL, { &centerX }, "pore axis, x coordinate [nm]" },
        { "-centerY", FALSE, etREAL, { &centerY }, "pore axis, y coordinate [nm]" },
        { "-dim", FALSE, etINT, { &dim }, "pore axis direction: 0 = x, 1 = y, 2 = z" },
        { "-dr", FALSE, etREAL, { &dr }, "radial bin width [nm]" },
        { "-Cr", FALSE, etREAL, { &Cr }, "inner radius of the histogram [nm]" },
        { "-CR", FALSE, etREAL, { &CR }, "outer radius of the histogram [nm]" },
    };
    hd.fnm = { { efXVG, "-o", "rdf_time_in_pore", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const int nbin = (CR - Cr) / dr;   // float arithmetic, :53
    itp::vecd volumn(nbin);
    volumn.fill(0);
    for (size_t i = 0; i != (size_t)nbin; ++i) volumn[i] = itp::pi * (std::pow(Cr + dr * (i + 1), 2) - std::pow(Cr + dr * i, 2)) * (z2 - z1);

    b2a_region_spec spec{};
    spec.grp = 0; spec.com = B2A_COM_MASS; spec.nfilter = 3;
    spec.fdim[0] = XX; spec.flow[0] = x1; spec.fup[0] = x2;
    spec.fdim[1] = YY; spec.flow[1] = y1; spec.fup[1] = y2;
    spec.fdim[2] = ZZ; spec.flow[2] = z1; spec.fup[2] = z2;
    spec.cx = centerX; spec.cy = centerY;                                                   // axis of the radial coordinate (no cylinder test)
    spec.nbdim = 1; spec.bdim[0] = 3; spec.blow[0] = Cr; spec.dbin[0] = dr; spec.nb[0] = nbin;   // meZ = (r - Cr) / dr, :91
    spec.per_frame = 1;
    const int acc = hd.b2aRegion(spec);

    auto file = hd.openWrite(hd.get_opt2fn("-o"));
    fmt::print(file, "{:8.4f} ", 0.0);
    for (int i = 0; i != nbin; ++i) fmt::print(file, "{:8.5f} ", Cr + i * dr + dr / 2);
    fmt::print(file, "\n");
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());
    std::vector<uint64_t> per;
    std::vector<float>    times;
    const size_t          nfr = hd.b2aSeries(acc, per, nullptr, &times);   // every frame, in trajectory order, after the loop
    for (size_t k = 0; k < nfr; ++k)
    {
        fmt::print(file, "{:8.4f} ", times[k]);
        for (int i = 0; i != nbin; ++i) fmt::print(file, "{:8.5f} ", (double)per[k * nbin + i] / volumn[i]);
        fmt::print(file, "\n");
    }
    fclose(file);
    return 0;
}
