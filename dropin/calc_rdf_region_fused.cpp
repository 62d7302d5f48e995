// calc_rdf_region, ported to the fused device accumulator (full-speed path).
//
// Same command line, same output file as the reference program src/general/calc_rdf_region.cpp; the per-frame
// loadPositionCenter + O(N^2) host loop (:70-95) is replaced by one b2aAccumulate() per resident K-frame batch,
// and the post-processing (:100-135) by b2a_rdf_normalise + the same print statements.  Compare with the
// UNMODIFIED reference program built against the same header (build/dropin/calc_rdf_region): that one is the
// compatibility mode, this one is what a maintainer writes to get the roofline numbers.
#include <itp/gmx>

#include <chrono>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);
    hd.ngrps = 2;

    int  nbin = 5, dim = 0, dim2 = 2;
    real lowPos = -100, upPos = 100, lowPos2 = -100, upPos2 = 100, Rm = 0;
    hd.pa = {
        { "-nbin", FALSE, etINT, { &nbin }, "number of histogram bins" },
        { "-up", FALSE, etREAL, { &upPos }, "upper bound of the selection slab [nm]" },
        { "-low", FALSE, etREAL, { &lowPos }, "lower bound of the selection slab [nm]" },
        { "-dim", FALSE, etINT, { &dim }, "axis the two bounds refer to: 0 = x, 1 = y, 2 = z" },
        { "-dim2", FALSE, etINT, { &dim2 }, "axis of the second slab (0 = x, 1 = y, 2 = z)" },
        { "-low2", FALSE, etREAL, { &lowPos2 }, "lower z bound [nm]" },
        { "-up2", FALSE, etREAL, { &upPos2 }, "upper z bound [nm]" },
        { "-Rm", FALSE, etREAL, { &Rm }, "largest pair distance [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "rdf_region", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    if (Rm < 1e-6)   // reference :52-57
    {
        Rm = std::min(hd.Lbox[0], hd.Lbox[1]);
        Rm = std::min(double(Rm), hd.Lbox[2]);
        Rm /= 2;
    }
    const int Rbin = Rm / 0.002;   // :58

    b2a_rdf_spec spec{};
    spec.grp0 = 0; spec.grp1 = 1; spec.com0 = spec.com1 = B2A_COM_MASS;
    spec.nbin = nbin; spec.low = lowPos; spec.up = upPos; spec.dim = dim;
    spec.dim2 = dim2; spec.low2 = lowPos2; spec.up2 = upPos2; spec.Rm = Rm; spec.Rbin = Rbin;
    const int acc = hd.b2aRdf(spec);

    const auto t0 = std::chrono::steady_clock::now();
    do
    {
        hd.b2aAccumulate(acc);        // every frame of the resident batch, on the device
    } while (hd.b2aReadNextBatch());

    std::vector<uint64_t> counts;
    uint64_t              stats[B2A_STAT_NSLOTS];
    hd.b2aRead(acc, counts, stats);   // synchronises: everything queued above has finished
    const double loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    itp::matd rdf(nbin, Rbin);
    b2a_rdf_normalise(counts.data(), nbin, Rbin, rdf.data());   // :100-115

    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    fmt::print(ofile, "# {} ", char('x' + dim));
    for (int i = 0; i != nbin + 1; ++i) fmt::print(ofile, "{:8.4f} ", lowPos + i * (upPos - lowPos) / nbin);
    fmt::print(ofile, "\n");
    for (int j = 0; j != Rbin; ++j)
    {
        fmt::print(ofile, "{:8.4f} ", j * 0.002);
        for (int i = 0; i != nbin; ++i) fmt::print(ofile, "{:8.4f} ", rdf(i, j));
        fmt::print(ofile, "\n");
    }
    fclose(ofile);
    fmt::print("frames {}  pair distances evaluated {}  re-evaluated in FP64 {}  out-of-bounds in the reference {}\n",
               stats[B2A_STAT_FRAMES], stats[B2A_STAT_PAIRS], stats[B2A_STAT_EXACT], stats[B2A_STAT_DROPPED]);
    // the frame loop alone (decode of every batch but the first, uploads, kernels, final read-back): what an ingestion benchmark wants
    fmt::print("frame loop {:.6f} s  = {:.1f} frames/s\n", loop_s, stats[B2A_STAT_FRAMES] / loop_s);
    return 0;
}
