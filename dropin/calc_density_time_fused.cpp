// calc_density_time (reference src/mofs/calc_density_time.cpp), ported to the device.
//
// Same command line and output file: the number of molecules inside the pore (z slab + cylinder) for every frame.  The per-frame
// loop (:38-54) becomes one b2aAccumulate() of a region accumulator with per-frame counters per resident batch; the series and its
// time stamps are fetched once after the loop (b2aSeries), in trajectory order whichever GPU processed a batch.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);

    real lowPos = 7.761, upPos = 12.239, centerX = 2.829, centerY = 2.056, Rmof = 0.927;
    hd.pa = {
        { "-upPos", FALSE, etREAL, { &upPos }, "upper bound of the pore region [nm]" },
        { "-lowPos", FALSE, etREAL, { &lowPos }, "lower bound of the pore region [nm]" },
        { "-centerX", FALSE, etREAL, { &centerX }, "pore axis, x coordinate [nm]" },
        { "-centerY", FALSE, etREAL, { &centerY }, "pore axis, y coordinate [nm]" },
        { "-Rmof", FALSE, etREAL, { &Rmof }, "pore radius [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "DensityTime", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    b2a_region_spec spec{};
    spec.grp = 0; spec.com = B2A_COM_MASS;
    spec.nfilter = 1; spec.fdim[0] = ZZ; spec.flow[0] = lowPos; spec.fup[0] = upPos; spec.fincl[0] = 0;
    spec.cyl = 1; spec.cx = centerX; spec.cy = centerY; spec.R2 = double(Rmof) * Rmof;   // :33
    spec.nbdim = 0; spec.per_frame = 1;
    const int acc = hd.b2aRegion(spec);

    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    // nothing is read back inside the loop (with several GPUs -- B2A_DEVICES -- that would serialise them): the per-frame counters
    // stay on the device that computed them and come back after the loop, ordered by frame
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());
    std::vector<uint64_t> per;
    std::vector<float>    times;
    const size_t          nfr = hd.b2aSeries(acc, per, nullptr, &times);
    for (size_t k = 0; k < nfr; ++k) fmt::print(ofile, "{:8.4f} {:10d}\n", times[k], (int)per[k]);
    fclose(ofile);
    return 0;
}
