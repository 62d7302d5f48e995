// calc_pair_dist_bulk (reference src/mofs/calc_pair_dist_bulk.cpp), ported to the fused device accumulator.
//
// Same command line and output file.  The O(N0 x N1) coordination-number loop per frame (:66-88) runs on the device (K5); the
// reference normalises every frame by its own number of selected molecules (totPair[k] += tmpPair[k] / numA, :89-92), so the
// per-frame tables of each batch are read back and folded in frame order (b2a_coord_fold does that arithmetic).
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);
    hd.ngrps = 2;

    int  NCM = 0, dim = 2;
    real Rmin = 0.8, lowPos = 0.0, upPos = 12.0;
    hd.pa = {
        { "-ncm", FALSE, etINT, { &NCM }, "molecular centre: 0 = mass, 1 = geometric, 2 = charge" },
        { "-Rmin", FALSE, etREAL, { &Rmin }, "first-shell cut-off: position of the first RDF minimum [nm]" },
        { "-dim", FALSE, etINT, { &dim }, "profile axis: 0 = x, 1 = y, 2 = z" },
        { "-low", FALSE, etREAL, { &lowPos }, "lower bound [nm]" },
        { "-up", FALSE, etREAL, { &upPos }, "upper bound [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "pair_dist", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const int      max_count = 4096;   // pairNumber values the table can hold (more neighbours than that are reported as dropped)
    b2a_coord_spec spec{};
    spec.grp0 = 0; spec.grp1 = 1; spec.com0 = spec.com1 = NCM; spec.Rmin = Rmin; spec.dim = dim; spec.low = lowPos; spec.up = upPos;
    spec.use_cyl = 0; spec.max_count = max_count;
    const int acc = hd.b2aCoord(spec);

    std::vector<double> totPair(max_count, 0.0);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());
    hd.b2aCoordFoldSeries(acc, max_count, totPair);   // per-frame tables of every frame, folded in frame order after the loop

    double totol = 0;
    for (double p : totPair) totol += p;       // :95-99 (absent keys contribute nothing)
    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    for (int k = 0; k < max_count; ++k)
        if (totPair[k] > 0) fmt::print(ofile, "{:8d}\t{:12.5f}\n", k, totPair[k] / totol);
    fclose(ofile);
    return 0;
}
