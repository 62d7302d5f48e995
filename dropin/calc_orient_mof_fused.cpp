// calc_orient_mof (reference src/mofs/calc_orient_mof.cpp), ported to the fused device accumulator.
//
// Same command line and output file.  The frame loop of the reference (:150-200: loadPosition, loadPositionCenter, getPoint,
// cylinder / slab filter, calcVector, calcAngle, two histogram updates per molecule) becomes one b2aAccumulate() per resident
// batch (K4); the post-processing (:201-215) is the reference's arithmetic on the returned integer counts.
#include <itp/gmx>

gmx_main(temp)
{
    itp::GmxHandle hd(argc, argv);

    rvec dbin = { 0.01, 0.01, 0.01 }, low = { 0.0, 0.0, 0.0 }, up = { 100, 100, 100 };
    real cylCenter1 = 0.0, cylCenter2 = 0.0, dr = 0.01, CR = 0.5, Cr = 0.0;
    int  nAngle = 90, NCM = 0, DIMN = 2, tp1 = 1, tp2 = 2, tp3 = 3, DIMNOrient = 2, method = 1;
    hd.pa = {
        { "-NCM", FALSE, etINT, { &NCM }, "molecular centre: 0 = mass, 1 = geometric, 2 = charge" },
        { "-dim", FALSE, etINT, { &DIMN }, "axis of the slab filter; the radius is taken in the other two dimensions" },
        { "-dbin", FALSE, etRVEC, { dbin }, "bin width [nm]" },
        { "-low", FALSE, etRVEC, { low }, "lower bound of the region [nm]" },
        { "-up", FALSE, etRVEC, { up }, "upper bound of the region [nm]" },
        { "-c1", FALSE, etREAL, { &cylCenter1 }, "circle centre, first coordinate [nm]" },
        { "-c2", FALSE, etREAL, { &cylCenter2 }, "circle centre, second coordinate [nm]" },
        { "-dr", FALSE, etREAL, { &dr }, "radial bin width [nm]" },
        { "-CR", FALSE, etREAL, { &CR }, "outer radius of the radial range [nm]" },
        { "-Cr", FALSE, etREAL, { &Cr }, "inner radius of the radial range [nm]" },
        { "-nA", FALSE, etINT, { &nAngle }, "number of angle bins" },
        { "-tp1", FALSE, etINT, { &tp1 }, "position of vector atom 1 inside the molecule" },
        { "-tp2", FALSE, etINT, { &tp2 }, "position of vector atom 2 inside the molecule" },
        { "-tp3", FALSE, etINT, { &tp3 }, "position of vector atom 3 inside the molecule" },
        { "-DIMNOrient", FALSE, etINT, { &DIMNOrient }, "axis of the wall vector" },
        { "-method", FALSE, etINT, { &method }, "molecular vector: 1 = sum of the two bond vectors, 2 = their cross product" }
    };
    hd.fnm = { { efXVG, "-o", "orient", ffWRITE } };
    hd.init();
    hd.readFirstFrame();

    const int    nbin1  = (CR - Cr) / dr;      // float arithmetic, :133
    const double dAngle = 180 / nAngle;        // integer division, :143

    b2a_orient_spec spec{};
    spec.grp = 0; spec.com = NCM; spec.DIMN = DIMN; spec.low = low[DIMN]; spec.up = up[DIMN];
    spec.c1 = cylCenter1; spec.c2 = cylCenter2; spec.dr = dr; spec.CR = CR; spec.Cr = Cr; spec.nAngle = nAngle;
    spec.tp1 = tp1; spec.tp2 = tp2; spec.tp3 = tp3; spec.DIMNOrient = DIMNOrient; spec.method = method;
    const int acc = hd.b2aOrient(spec);
    do
    {
        hd.b2aAccumulate(acc);
    } while (hd.b2aReadNextBatch());

    // counts: orient(meZ, meA) at meZ + meA * nbin1, then NCMDens[meZ] (include/b2a.h)
    std::vector<uint64_t> counts;
    hd.b2aRead(acc, counts);
    auto file = hd.openWrite(hd.get_opt2fn("-o"));
    for (int i = 0; i != nAngle; ++i)
    {
        fmt::print(file, "{:8.4e}  ", (i + 0.5) * dAngle);
        for (int j = 0; j != nbin1; ++j)
        {
            double    v   = (double)counts[(size_t)j + (size_t)i * nbin1];
            const int ncm = (int)counts[(size_t)nbin1 * nAngle + j];
            if (ncm > 0.0) v /= ncm;           // :207-210
            fmt::print(file, "{:8.4e} ", v);
        }
        fmt::print(file, "\n");
    }
    fclose(file);
    return 0;
}
