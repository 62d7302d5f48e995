// calc_numdens_pore (reference src/mofs/calc_numdens_pore.cpp), ported to the device.
//
// Same command line and output file.  bInRegion (:12-27: every centre component cast to `real`, moved into [0, L] by repeated
// -+L, tested against the box low <= p < up) and the counting loop (:72-80) become one b2aAccumulate() of a region accumulator
// (`wrap`, three filters, no bins) per resident batch.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);

    rvec lowPos = { 0, 0, 0 }, upPos = { 30, 30, 30 };
    int  com = 0;
    hd.pa = {
        { "-low", FALSE, etRVEC, { lowPos }, "lower bound" },
        { "-up", FALSE, etRVEC, { upPos }, "upper bound" },
        { "-com", FALSE, etINT, { &com }, "molecular centre: 0 = mass, 1 = geometric, 2 = charge" }
    };
    hd.fnm = { { efXVG, "-o", "calc_numdens_pore", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    b2a_region_spec spec{};
    spec.grp = 0; spec.com = com; spec.wrap = 1; spec.nfilter = 3; spec.nbdim = 0;
    for (int m = 0; m < 3; ++m) { spec.fdim[m] = m; spec.flow[m] = lowPos[m]; spec.fup[m] = upPos[m]; }
    const int acc = hd.b2aRegion(spec);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    double numdens = (double)counts[0];
    numdens /= hd.nframe;       // :82

    auto file = hd.openWrite(hd.get_ftp2fn(efXVG));
    fmt::print(file, "{:8.4f}\n", numdens);
    fclose(file);
    return 0;
}
