// calc_pair_dist (reference src/mofs/calc_pair_dist.cpp), ported to the fused device accumulator.
//
// Same command line and output file.  Like calc_pair_dist_bulk, with the additional selection of the central molecules by their
// distance from the pore axis (:85-86); the reference loads mass centres whatever -ncm says (:78-79), and so does this port.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);
    hd.ngrps = 2;

    int  dim = 2, NCM = 0;
    real lowPos = 7.073, upPos = 12.927, centerX = 2.829, centerY = 2.056, Rmin = 0.8, Rlow = 0, Rup = 0.4;
    hd.pa = {
        { "-dim", FALSE, etINT, { &dim }, "axis of the slab: 0 = x, 1 = y, 2 = z" },
        { "-up", FALSE, etREAL, { &upPos }, "upper bound of the selection slab [nm]" },
        { "-low", FALSE, etREAL, { &lowPos }, "lower bound of the selection slab [nm]" },
        { "-centerX", FALSE, etREAL, { &centerX }, "axis position, x [nm]" },
        { "-centerY", FALSE, etREAL, { &centerY }, "axis position, y [nm]" },
        { "-ncm", FALSE, etINT, { &NCM }, "molecular centre: 0 = mass, 1 = geometric, 2 = charge" },
        { "-Rmin", FALSE, etREAL, { &Rmin }, "first-shell cut-off: position of the first RDF minimum [nm]" },
        { "-Rlow", FALSE, etREAL, { &Rlow }, "smallest distance from the pore axis [nm]" },
        { "-Rup", FALSE, etREAL, { &Rup }, "largest distance from the pore axis [nm]" }
    };
    hd.fnm = { { efXVG, "-o", "pair_dist", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const int      max_count = 4096;
    b2a_coord_spec spec{};
    spec.grp0 = 0; spec.grp1 = 1; spec.com0 = spec.com1 = B2A_COM_MASS; spec.Rmin = Rmin; spec.dim = dim; spec.low = lowPos; spec.up = upPos;
    spec.use_cyl = 1; spec.cx = centerX; spec.cy = centerY; spec.Rlow = Rlow; spec.Rup = Rup; spec.max_count = max_count;
    const int acc = hd.b2aCoord(spec);

    std::vector<double> totPair(max_count, 0.0);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());
    hd.b2aCoordFoldSeries(acc, max_count, totPair);   // per-frame tables of every frame, folded in frame order after the loop

    double totol = 0;
    for (double p : totPair) totol += p;
    auto ofile = hd.openWrite(hd.get_opt2fn("-o"));
    for (int k = 0; k < max_count; ++k)
        if (totPair[k] > 0) fmt::print(ofile, "{:8d}\t{:12.5f}\n", k, totPair[k] / totol);
    fclose(ofile);
    return 0;
}
