// The README's density template (reference README.md:24-177, BASELINE config 1), ported to the fused device accumulator.
//
// Same command line and output file as the template; the per-frame loadPositionCenter + binning loop (README.md:129-150)
// becomes one b2aAccumulate() per resident K-frame batch (K2: centres are reduced and binned in one pass over the float
// coordinates, nothing is written back but the histogram).  The unmodified template built against the same header
// (build/dropin/readme_template) is the compatibility mode; both write identical files.
#include <itp/gmx>

#include <chrono>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);
    hd.ngrps = 1;
    const int grp = 0;

    int  nbin = 100, dim = 2;
    real lowPos = 0, upPos = 30;
    hd.pa = {
        { "-nbin", FALSE, etINT, { &nbin }, "number of histogram bins" },
        { "-dim", FALSE, etINT, { &dim }, "axis of the slab: 0 = x, 1 = y, 2 = z" },
        { "-up", FALSE, etREAL, { &upPos }, "upper bound of the selection slab [nm]" },
        { "-low", FALSE, etREAL, { &lowPos }, "lower bound of the selection slab [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "analysisOutput", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const double dbin = (upPos - lowPos) / nbin;   // float arithmetic widened, exactly as README.md:127

    b2a_density_spec spec{};
    spec.grp = grp; spec.com = B2A_COM_MASS; spec.dim = dim; spec.low = lowPos; spec.up = upPos; spec.dbin = dbin; spec.nbin = nbin;
    spec.upper_inclusive = 0;   // lowPos <= p < upPos (README.md:145)
    spec.dim2 = -1;
    const int  acc = hd.b2aDensity(spec);
    const auto t0  = std::chrono::steady_clock::now();
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    const double loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    itp::vecd density(nbin);
    for (int i = 0; i < nbin; i++) density[i] = (double)counts[i];
    density /= hd.nframe;       // README.md:154

    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    for (int i = 0; i < nbin; i++) fmt::print(ofile, "{:8.5f} {:10.5f}\n", lowPos + dbin / 2 + dbin * i, density[i]);
    fclose(ofile);
    // the frame loop alone (decode of every batch but the first, uploads, kernels, final read-back), for ingestion benchmarks
    fmt::print("frame loop {:.6f} s  = {:.1f} frames/s\n", loop_s, hd.nframe / loop_s);
    return 0;
}
