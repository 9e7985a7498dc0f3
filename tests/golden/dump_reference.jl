# dump_reference.jl -- run the UNMODIFIED reference (Clpr/AdaptiveSG.jl, Julia >= 1.9) and dump what pins parity:
#   julia --project=/path/to/AdaptiveSG.jl -t 1 tests/golden/dump_reference.jl [outdir = tests/golden]
# Writes julia_cfg1.bin / julia_cfg2.bin (format below, little endian), read by tests/test_reference_dump.py, which
# then checks the oracle AND the GPU path against them (node set bit-exact, values within 1e-12 / 1e-14).
# Not executed in this repository's environment (no Julia in the image): commit the two files it writes.
#   int64  D, N, M, nthreads | float64 wall_train, wall_adapt, wall_eval | float64 lb[D], ub[D]
#   int64  levels[N*D] (row-major), indices[N*D], depths[N] | float64 fvals[N], coefs[N]
#   float64 X[M*D] (row-major, domain units), y[M] = G.(points)
import AdaptiveSG as asg
using Random
function dump(name, D, f, tr, ad, M, outdir)
    G = asg.SparseGridInterpolant(D)
    t0 = time(); asg.train!(G, f, tr...; verbose = false); t1 = time()
    ad === nothing || asg.adapt!(G, f, ad[1]; verbose = false, tol = ad[2], toltype = :absolute, parallel = false)
    t2 = time()
    S = asg.serialize(G)                                           # src/io.jl:25-41
    X = rand(MersenneTwister(42), M, D)
    t3 = time(); y = [G(X[m, :]) for m in 1:M]; t4 = time()       # src/class.jl:747-761
    open(joinpath(outdir, "julia_$name.bin"), "w") do io
        write(io, Int64[D, length(G), M, Threads.nthreads()], Float64[t1 - t0, t2 - t1, t4 - t3])
        write(io, Float64.(S[:lb]), Float64.(S[:ub]))
        write(io, Int64.(permutedims(S[:levels])), Int64.(permutedims(S[:indices])), Int64.(S[:depths]))
        write(io, S[:fvals], S[:coefs], permutedims(X), y)
    end
    println("$name: N = $(length(G)), threads = $(Threads.nthreads()), train $(t1-t0) s, adapt $(t2-t1) s, ",
            "eval of $M points $(t4-t3) s")
end
outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__)
dump("cfg1", 3, X -> sum(sin.(X .* 2π)), (7, 5), nothing, 10_000, outdir)            # README.md:40-50
dump("cfg2", 2, X -> -1 / (0.1 + sum(X)), (4, (3, 8)), (20, 1e-5), 10_000, outdir)   # README.md:194-210
