# AdaptiveSGCUDA.jl -- thin `ccall` shim that keeps AdaptiveSG.jl's API a drop-in while the hot
# path (evaluation, basis rows, node coordinates, the surplus step of train!/adapt!) runs in
# libasg_cuda.so on a B200.  No CUDA.jl, no kernels in Julia, no CPU fallback.
#
# NOT EXECUTED IN THIS REPOSITORY'S ENVIRONMENT: Julia is not installed in the build image or on
# the GPU boxes (SURVEY.md, "Read this first").  The Python mirror (../interpolant.py) implements
# the same host logic over the same C ABI and is what the test-suite drives.  Every method below
# names the reference method it replaces (paths relative to the AdaptiveSG.jl checkout).
#
# Usage:   ENV["ASG_CUDA_LIB"] = "/path/to/libasg_cuda.so";  import AdaptiveSGCUDA as asg
module AdaptiveSGCUDA

using SparseArrays, Printf

export SparseGridInterpolant, rsg!, train!, adapt!, coefficients, basis, dehierarchization_matrix,
       hierarchization_matrix, hierarchize!, extrapolate, serialize, deserialize, maxlevels, sync!
# every export of the reference's math.jl (src/math.jl:9-22), so that the module can stand in for the package
export nodedepth, perturb, scale, ϕ, isvalid, get_x, isboundary, boundaryflag, allnodes1d, allindices_of_level, child,
       parent, ghost_neighbor

const LIB = get(ENV, "ASG_CUDA_LIB", "libasg_cuda")

struct AsgError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::AsgError) = print(io, "libasg_cuda error ", e.code, ": ", e.msg)

last_error() = unsafe_string(ccall((:asg_last_error, LIB), Cstring, ()))
function check(rc::Int32)
    rc == 0 && return nothing
    msg = last_error()
    rc == -2 && throw(ArgumentError(msg))      # ASG_EINVALID_NODE: where the reference throws (src/math.jl:355, 474)
    throw(AsgError(rc, msg))
end

const _initialised = Ref(false)
function init(device::Integer = parse(Int, get(ENV, "ASG_CUDA_DEVICE", "0")))
    check(ccall((:asg_init, LIB), Int32, (Int32,), device))
    _initialised[] = true
end
"""
    init_devices(devices)

One Julia process drives several GPUs of the box (`asg_init_devices`): every grid keeps a replica of its packed
layouts per device and `evaluate` / `surplus` / `train!` / `hierarchize!` split their points over the devices inside
the library -- no Distributed.jl, no MPI, no NCCL. `devices[1]` is the primary device. This is what stands in for the
reference's only parallel entry, `adapt!(...; parallel = true)` (src/train.jl:332-340). `ENV["ASG_CUDA_DEVICES"] =
"0,1,2,3"` makes the first constructor call do it.
"""
function init_devices(devices::AbstractVector{<:Integer})
    d = Int32.(collect(devices))
    check(ccall((:asg_init_devices, LIB), Int32, (Ptr{Int32}, Int32), d, length(d)))
    _initialised[] = true
end
_auto_init() = haskey(ENV, "ASG_CUDA_DEVICES") ? init_devices(parse.(Int, split(ENV["ASG_CUDA_DEVICES"], ","))) : init()

# ---------------------------------------------------------------------------------------------
# node store: same public fields as src/class.jl:281-293, plus the device handle
# ---------------------------------------------------------------------------------------------
mutable struct SparseGridInterpolant{D}
    lb::NTuple{D,Float64}
    ub::NTuple{D,Float64}
    levels::Matrix{Int}
    indices::Matrix{Int}
    depths::Vector{Int}
    fvals::Vector{Float64}
    coefs::Vector{Float64}
    handle::Ptr{Cvoid}     # asg_grid*
    synced::Bool           # device mirror matches the arrays above

    function SparseGridInterpolant(xdim::Int; lb = zeros(Float64, xdim), ub = ones(Float64, xdim))
        # same asserts, same messages as src/class.jl:301-306
        @assert xdim > 0 "Dimension must be positive."
        @assert length(lb) == xdim "Lower bound must match dimension."
        @assert length(ub) == xdim "Upper bound must match dimension."
        @assert all(ub .> lb) "Upper bound must be greater than lower bound."
        @assert all(isfinite, lb) "Lower bound must be finite."
        @assert all(isfinite, ub) "Upper bound must be finite."
        _initialised[] || _auto_init()
        lbv = collect(Float64, lb); ubv = collect(Float64, ub)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:asg_grid_create, LIB), Int32, (Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}),
                    xdim, lbv, ubv, h))
        G = new{xdim}(NTuple{xdim,Float64}(lbv), NTuple{xdim,Float64}(ubv), Matrix{Int}(undef, 0, xdim),
                      Matrix{Int}(undef, 0, xdim), Int[], Float64[], Float64[], h[], false)
        # asg_grid_destroy is safe from any thread and does not block (finalizers run anywhere)
        finalizer(g -> ccall((:asg_grid_destroy, LIB), Int32, (Ptr{Cvoid},), g.handle), G)
        return G
    end
end

# users assign fields directly (README.md:237-241, src/io.jl:64-69): any re-assignment of a node
# array invalidates the device mirror; in-place edits need an explicit sync!(G)
function Base.setproperty!(G::SparseGridInterpolant, f::Symbol, v)
    f in (:levels, :indices, :depths, :coefs) && setfield!(G, :synced, false)
    setfield!(G, f, convert(fieldtype(typeof(G), f), v))
end

"""Push the Julia arrays to the device (full replace). Column-major N x D goes over as is:
row stride 1, column stride N -- no transpose, no copy on the Julia side."""
function sync!(G::SparseGridInterpolant{D}) where {D}
    N = size(G.levels, 1)
    @assert size(G.indices) == (N, D) && length(G.coefs) == N "inconsistent node arrays"
    GC.@preserve G check(ccall((:asg_grid_upload, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Ptr{Float64}),
        G.handle, N, G.levels, G.indices, 1, max(N, 1), G.coefs))
    setfield!(G, :synced, true)
    return G
end
ensure!(G) = (getfield(G, :synced) || sync!(G); G.handle)

# Base overloads kept verbatim in behaviour (src/class.jl:321-366)
Base.length(G::SparseGridInterpolant) = length(G.fvals)
Base.ndims(::SparseGridInterpolant{D}) where {D} = D
Base.size(G::SparseGridInterpolant{D}) where {D} = (length(G.fvals), D)
Base.in(x::AbstractVector, G::SparseGridInterpolant) = all(x .>= G.lb) && all(x .<= G.ub)
Base.clamp(x::AbstractVector, G::SparseGridInterpolant) = clamp.(x, G.lb, G.ub)
Base.rand(G::SparseGridInterpolant{D}) where {D} = rand(Float64, D) .* (G.ub .- G.lb) .+ G.lb
Base.rand(G::SparseGridInterpolant{D}, n::Int) where {D} = rand(Float64, D, n) .* (G.ub .- G.lb) .+ G.lb
Base.LinRange(G::SparseGridInterpolant, d::Int, n::Int) = LinRange(G.lb[d], G.ub[d], n)
function Base.show(io::IO, G::SparseGridInterpolant{D}) where {D}
    @printf(io, "SparseGridInterpolant f(x): R^%d --> R\n#nodes: %d\nDomain:\n", D, length(G))
    for d in 1:D; @printf(io, "  x[%d] in [%.2f, %.2f]\n", d, G.lb[d], G.ub[d]); end
end
maxlevels(G::SparseGridInterpolant{D}) where {D} = isempty(G.levels) ? zeros(Int, D) : vec(maximum(G.levels, dims = 1))
coefficients(G::SparseGridInterpolant) = G.coefs                        # src/class.jl:414-416

# host-side integer helpers the refinement logic needs (src/math.jl:59-61, 103-108, 418-448, 651-661, 907-985)
power2(x::Int) = 2 << (x - 1)
nodedepth(ls) = sum(ls) - length(ls) + 1
scale(x::AbstractVector, flb, fub; to_lb = 0.0, to_ub = 1.0) = to_lb .+ (to_ub .- to_lb) .* (x .- flb) ./ (fub .- flb)
function allindices_of_level(l::Int)
    l > 2 && return collect(1:2:(power2(l - 1) - 1))
    l == 2 && return [0, 2]
    l == 1 && return [1]
    throw(ArgumentError("invalid level $l"))
end
function isvalid(Ls::AbstractVector{Int}, Is::AbstractVector{Int})
    (length(Ls) < 1 || length(Is) != length(Ls)) && return false
    for (l, i) in zip(Ls, Is)
        ok = l == 1 ? i == 1 : l == 2 ? (i == 0 || i == 2) : (l > 2 && i >= 1 && isodd(i))
        ok || return false
    end
    return nodedepth(Ls) >= 1
end
function child1(l::Int, i::Int, right::Bool)
    l > 2 && return (l + 1, right ? 2i + 1 : 2i - 1)
    l == 2 && i == 0 && return (3, right ? 1 : -1)
    l == 2 && i == 2 && return (3, right ? -1 : 3)
    l == 1 && return (2, right ? 2 : 0)
    throw(ArgumentError("invalid level-index pair ($l, $i)"))
end
function child(Ls::AbstractVector{Int}, Is::AbstractVector{Int}, dims::Int; side::Symbol = :left)
    side in (:left, :right) || throw(ArgumentError("side must be :left or :right"))
    L = collect(Int, Ls); I = collect(Int, Is)
    L[dims], I[dims] = child1(L[dims], I[dims], side == :right)
    return (L, I)
end

# the basis function itself (src/math.jl:332-391): evaluated by the library (asg_phi), bit-identical to the reference's
# operation order; an invalid (l, i) pair throws ArgumentError like src/math.jl:355
function _phi(x::Vector{Float64}, L, I, n::Int, D::Int)
    _initialised[] || _auto_init()
    out = Vector{Float64}(undef, n)
    GC.@preserve x L I out check(ccall((:asg_phi, LIB), Int32, (Int64, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                       n, D, x, L === nothing ? C_NULL : L, I === nothing ? C_NULL : I, out))
    return out
end
ϕ(x::Float64)::Float64 = _phi([x], nothing, nothing, 1, 1)[1]                                       # src/math.jl:332-334
ϕ(x::Float64, l::Int, i::Int)::Float64 = _phi([x], Int64[l], Int64[i], 1, 1)[1]                     # src/math.jl:345-357
function ϕ(x::AbstractVector, levels::AbstractVector, indices::AbstractVector; safe::Bool = false)::Float64  # 376-391
    D = length(x)
    if safe
        @assert length(levels) == D "levels must match the dimension of x."
        @assert length(indices) == D "indices must match the dimension of x."
    end
    return _phi(collect(Float64, x), collect(Int64, levels), collect(Int64, indices), 1, D)[1]
end

# node arithmetic the reference exports next to it (host-side integer helpers; src/math.jl:516-607, 621-639, 704-879,
# 1002-1024), written against their documented behaviour
perturb(Ls, Is, dims::Int, l::Int, i::Int) = (L = collect(Int, Ls); I = collect(Int, Is); L[dims] = l; I[dims] = i; (L, I))
isboundary(l::Int, i::Int) = l == 2 && (i == 0 || i == 2)
function boundaryflag(l::Int, i::Int)
    l == 2 || return 0
    i == 0 && return -1
    i == 2 && return 1
    throw(ArgumentError("invalid level-index pair ($l, $i)"))
end
function fraction2li(q::Rational)
    q == 0 // 1 && return (2, 0)
    q == 1 // 1 && return (2, 2)
    q == 1 // 2 && return (1, 1)
    return (trailing_zeros(q.den) + 1, Int(q.num))
end
function allnodes1d(l::Int)
    l < 1 && throw(ArgumentError("level must be >= 1 but got $l"))
    l == 1 && return ([1], [1])
    l == 2 && return ([1, 2, 2], [1, 0, 2])
    den = power2(l - 1)
    pairs = [fraction2li(k // den) for k in 0:den]
    return (first.(pairs), last.(pairs))
end
function parent(Ls::AbstractVector{Int}, Is::AbstractVector{Int}, dims::Int)
    l, i = Ls[dims], Is[dims]
    inew = l > 3 ? div(i, 4) * 2 + 1 : (l, i) == (3, 1) ? 0 : (l, i) == (3, 3) ? 2 :
           (l == 2 && (i == 0 || i == 2)) ? 1 : l == 1 ? 0 : throw(ArgumentError("invalid level-index pair ($l, $i)"))
    return perturb(Ls, Is, dims, l - 1, inew)
end
ghost_distance(lmax::Int) = 1.0 / power2(lmax - 1)
function ghost_neighbor(Ls::AbstractVector{Int}, Is::AbstractVector{Int}, dims::Int, l0::Int, side::Symbol = :left)
    side in (:left, :right) || throw(ArgumentError("side must be :left or :right"))
    l0 < 1 && throw(ArgumentError("invalid level $l0"))
    lj, ij = Ls[dims], Is[dims]; step = side == :left ? -1 : 1; none = (0, -1)
    if l0 == 1
        new = none
    elseif l0 == 2
        order = [(2, 0), (1, 1), (2, 2)]; k = findfirst(==((lj, ij)), order)
        new = (k === nothing || !(1 <= k + step <= 3)) ? none : order[k + step]
    else
        lj < 1 && throw(ArgumentError("lj<1, invalid input node"))
        den = power2(l0 - 1)
        pos = lj == 1 ? (ij == 1 ? den ÷ 2 : throw(ArgumentError("invalid input node"))) :
              lj == 2 ? (ij == 0 ? 0 : ij == 2 ? den : throw(ArgumentError("invalid input node"))) :
              lj <= l0 ? ij * power2(l0 - lj) : throw(ArgumentError("lj > l0 found, no grid defined"))
        k = pos + step
        new = (k < 0 || k > den) ? none : (lj == 2 && 0 < k < den) ? (l0, k) : fraction2li(k // den)
    end
    return perturb(Ls, Is, dims, new[1], new[2])
end
get_x(l::Int, i::Int)::Float64 = l > 2 && isodd(i) ? Float64(i / power2(l - 1)) : (l, i) == (2, 0) ? 0.0 : (l, i) == (2, 2) ? 1.0 :
                                 (l, i) == (1, 1) ? 0.5 : throw(ArgumentError("invalid level-index pair ($l, $i)"))

# ---------------------------------------------------------------------------------------------
# evaluation: replaces _interpolate / _interpolate_if / _interpolate_until_level / G(x)
# (src/class.jl:629-680, 747-761; call sites src/class.jl:756,759)
# ---------------------------------------------------------------------------------------------
"""
    evaluate(G, X; dims = 1, max_depth = -1)

Batched `[G(x) for x in eachrow(X)]` (`dims = 1`) or `eachcol` (`dims = 2`, the layout `rand(G, n)`
returns). Strides are passed through, nothing is transposed. `max_depth >= 0` applies the mask
of `_interpolate_until_level`.
"""
function evaluate(G::SparseGridInterpolant{D}, X::AbstractMatrix{Float64}; dims::Int = 1, max_depth::Int = -1) where {D}
    M, sp, sd = dims == 1 ? (size(X, 1), stride(X, 1), stride(X, 2)) : (size(X, 2), stride(X, 2), stride(X, 1))
    @assert size(X, dims == 1 ? 2 : 1) == D "Input point must match the dimension of the grid."
    out = Vector{Float64}(undef, M)
    h = ensure!(G)
    GC.@preserve X out check(ccall((:asg_eval, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}), h, X, M, sp, sd, max_depth, out))
    return out
end
(G::SparseGridInterpolant)(X::AbstractMatrix) = evaluate(G, convert(Matrix{Float64}, X))   # additive: batched rows

"""
    extrapolate(G, X)

Batched linear extrapolation of `G` to the rows of `X` (documented intent of `_extrapolate`, src/class.jl:695-737,
which the reference's call site never reaches): one batched evaluation at the clamped points plus one per dimension
for the points outside in that dimension. Unit-cube lengths; slopes against the value at the boundary point.
"""
function extrapolate(G::SparseGridInterpolant{D}, X::AbstractMatrix{Float64}) where {D}
    w = G.ub .- G.lb
    x01 = (X .- G.lb') ./ w'
    xb01 = clamp.(x01, 0.0, 1.0)
    to_x(u) = ifelse.(u .>= 1.0, G.ub', ifelse.(u .<= 0.0, G.lb', G.lb' .+ w' .* u))   # faces map to lb / ub exactly
    f0 = evaluate(G, to_x(xb01))
    fval = copy(f0)
    maxl = maxlevels(G)
    for j in 1:D
        out = findall(.!((x01[:, j] .>= 0.0) .& (x01[:, j] .<= 1.0)))
        isempty(out) && continue
        h = 1.0 / (1 << (maxl[j] - 1))                          # ghost_distance, src/math.jl:888-890
        right = x01[out, j] .> 1.0
        xn = xb01[out, :]
        xn[:, j] .-= ifelse.(right, h, -h)
        fn = evaluate(G, to_x(xn))
        slope = ifelse.(right, f0[out] .- fn, fn .- f0[out]) ./ h
        fval[out] .+= slope .* (x01[out, j] .- xb01[out, j])
    end
    return fval
end

function (G::SparseGridInterpolant{D})(x::AbstractVector; safe::Bool = true, extrapolate::Bool = false) where {D}
    safe && @assert length(x) == D "Input point must match the dimension of the grid."
    # extrapolate = true is unreachable in the reference (undefined `__extrapolate`, src/class.jl:754)
    extrapolate && (x in G) && throw(UndefVarError(:__extrapolate))
    return evaluate(G, reshape(collect(Float64, x), 1, D))[1]
end

"alpha = fvals - G_{depth <= max_depth}(X): residual step of src/train.jl:176-183, 362-363, 381-382"
function surplus(G::SparseGridInterpolant{D}, X::Matrix{Float64}, fvals::Vector{Float64}; max_depth::Int = -1) where {D}
    M = size(X, 1); out = Vector{Float64}(undef, M)
    GC.@preserve X fvals out check(ccall((:asg_surplus, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
        ensure!(G), X, M, 1, max(M, 1), fvals, max_depth, out))
    return out
end

"get_x rows for arbitrary (levels, indices) in G's domain (src/math.jl:490-497), on the device"
function get_x(G::SparseGridInterpolant{D}, Ls::Matrix{Int}, Is::Matrix{Int}) where {D}
    n = size(Ls, 1); X = Matrix{Float64}(undef, n, D)
    GC.@preserve Ls Is X check(ccall((:asg_nodes_x_li, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Ptr{Float64}, Int64, Int64),
        G.handle, n, Ls, Is, 1, max(n, 1), X, 1, max(n, 1)))
    return X
end
Base.stack(G::SparseGridInterpolant{D}) where {D} =                      # src/class.jl:375-387
    isempty(G.levels) ? Matrix{Float64}(undef, 0, D) : get_x(G, G.levels, G.indices)

# ---------------------------------------------------------------------------------------------
# basis family (src/class.jl:434-567); documented intent, see DESIGN.md for defects B1/B2
# ---------------------------------------------------------------------------------------------
function basis(Xs::AbstractMatrix, G::SparseGridInterpolant{D}; safe::Bool = true, dropzeros::Bool = true,
               atol::Float64 = 1e-16) where {D}
    X = convert(Matrix{Float64}, Xs); M = size(X, 1); N = length(G); h = ensure!(G)
    nnzr = Vector{Int64}(undef, M)
    GC.@preserve X check(ccall((:asg_basis_count, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Float64, Int32, Ptr{Int64}),
        h, X, M, 1, max(M, 1), atol, dropzeros, nnzr))
    rowptr = vcat(0, cumsum(nnzr)); rows = Vector{Int64}(undef, rowptr[end]); val = Vector{Float64}(undef, rowptr[end])
    colptr = Vector{Int64}(undef, N + 1)
    # compressed sparse COLUMNS straight from the device (stable sort of the entries by column): no host transposition
    GC.@preserve X check(ccall((:asg_basis_fill_csc, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Float64, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        h, X, M, 1, max(M, 1), atol, dropzeros, rowptr, colptr, rows, val))
    return SparseMatrixCSC(M, N, colptr .+ 1, rows .+ 1, val)             # 0-based -> 1-based
end
basis(x::AbstractVector, G::SparseGridInterpolant; kw...) = sparsevec(vec(Matrix(basis(reshape(collect(Float64, x), 1, :), G; kw...))))
basis(G::SparseGridInterpolant{D}; kw...) where {D} =
    isempty(G.levels) ? SparseMatrixCSC{Float64,Int}(undef, 0, length(G)) : basis(stack(G), G; kw...)
dehierarchization_matrix(G::SparseGridInterpolant; kw...) = basis(G; kw...)
hierarchization_matrix(G::SparseGridInterpolant; kw...) =                 # src/class.jl:591-604 (host solve)
    (E = dehierarchization_matrix(G; kw...); sparse(E \ Matrix{Float64}(SparseArrays.I, size(E)...)))

"""
    hierarchize!(G, fvals = G.fvals)

`G.coefs = E \\ fvals` by forward substitution over depths (E is unit lower-triangular under a depth
sort): one fused surplus call per depth, no LU. Additive API (the reference offers only the dense
`hierarchization_matrix`, src/class.jl:591-604).
"""
function hierarchize!(G::SparseGridInterpolant, fvals::Vector{Float64} = G.fvals)
    h = ensure!(G)
    for lnow in 1:maximum(G.depths; init = 0)
        idx = findall(G.depths .== lnow); isempty(idx) && continue
        f = fvals[idx]; al = similar(f); idx0 = Int64.(idx .- 1)
        check(ccall((:asg_surplus_nodes, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}),
                    h, length(idx), idx0, f, lnow - 1, 1, al))
        G.coefs[idx] = al
    end
    return G.coefs
end

# ---------------------------------------------------------------------------------------------
# I/O (src/io.jl:9-72): same Dict schema; deserialize leaves the mirror unsynced
# ---------------------------------------------------------------------------------------------
# `packed = true` adds :packed, the device-format blob of the uploaded grid (asg_grid_export): deserialize then skips
# validation, sorting and packing (asg_grid_import); a blob of a build with other layout constants is ignored.
function serialize(G::SparseGridInterpolant{D}; packed::Bool = false) where {D}
    dat = Dict{Symbol,Any}(:D => D, :lb => collect(G.lb), :ub => collect(G.ub),
        :levels => G.levels, :indices => G.indices, :depths => G.depths, :fvals => G.fvals, :coefs => G.coefs)
    if packed
        h = ensure!(G); n = Ref{Int64}(0)
        check(ccall((:asg_grid_export, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}), h, C_NULL, 0, n))
        buf = Vector{UInt8}(undef, n[])
        GC.@preserve buf check(ccall((:asg_grid_export, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int64, Ref{Int64}), h, buf, n[], n))
        dat[:packed] = buf
    end
    return dat
end
function deserialize(dat::AbstractDict)
    for k in (:lb, :ub, :levels, :indices, :depths, :fvals, :coefs); @assert haskey(dat, k) "Serialized data must contain $k."; end
    G = SparseGridInterpolant(dat[:D], lb = dat[:lb], ub = dat[:ub])
    G.levels = dat[:levels]; G.indices = dat[:indices]; G.depths = dat[:depths]; G.fvals = dat[:fvals]; G.coefs = dat[:coefs]
    if haskey(dat, :packed)
        buf = dat[:packed]::Vector{UInt8}
        # the blob is installed only if it describes exactly these arrays (serialize aliases them: an in-place update
        # after the blob was written leaves it stale); -6 = ASG_EUNSUPPORTED (other build), -7 = ASG_ESTATE (stale /
        # foreign blob): both fall back to the normal upload on first use
        N = size(G.levels, 1)
        rc = GC.@preserve buf G ccall((:asg_grid_import_checked, LIB), Int32,
            (Ptr{Cvoid}, Ptr{UInt8}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Ptr{Float64}),
            G.handle, buf, length(buf), N, G.levels, G.indices, 1, max(N, 1), G.coefs)
        rc == 0 ? setfield!(G, :synced, true) : (rc in (-6, -7) || check(rc))
    end
    return G
end

# ---------------------------------------------------------------------------------------------
# training (src/train.jl): host logic unchanged in meaning, ONE kernel call per depth / level
# ---------------------------------------------------------------------------------------------
function rsg!(G::SparseGridInterpolant{D}, maxdepth::Int, maxlevels::Union{NTuple{D,Int},Int}) where {D}
    any(maxlevels .< 1) && throw(ArgumentError("maxlevels must be >= 1"))
    maxdepth < 2 && throw(ArgumentError("maxdepth must be >= 2"))
    ml = isa(maxlevels, NTuple) ? maxlevels : ntuple(_ -> maxlevels, D)
    Ls = NTuple{D,Int}[]; Is = NTuple{D,Int}[]
    # same visiting order as src/train.jl:56-63 (first dimension fastest), pruned instead of filtered
    function rec(d::Int, ks::Vector{Int}, remaining::Int)
        if d == 0
            kt = NTuple{D,Int}(ks)
            for it in Iterators.product((allindices_of_level(k) for k in kt)...); push!(Ls, kt); push!(Is, it); end
            return
        end
        for k in 1:min(ml[d], remaining - (d - 1)); ks[d] = k; rec(d - 1, ks, remaining - k); end
    end
    rec(D, ones(Int, D), maxdepth + D - 1)
    G.levels = permutedims(reinterpret(reshape, Int, Ls)); G.indices = permutedims(reinterpret(reshape, Int, Is))
    D == 1 && (G.levels = reshape(collect(first.(Ls)), :, 1); G.indices = reshape(collect(first.(Is)), :, 1))
    G.depths = [nodedepth(l) for l in Ls]
    G.fvals = zeros(length(Ls)); G.coefs = zeros(length(Ls))
    return nothing
end

function train!(G::SparseGridInterpolant{D}, f::Function, maxdepth::Int, maxlevels::Union{NTuple{D,Int},Int};
                verbose::Bool = false) where {D}
    rsg!(G, maxdepth, maxlevels)
    tic = time()
    verbose && println("Regular sparse grid training starts\n", "# dimensions: $D, # total nodes: $(length(G))")
    h = ensure!(G)
    for lnow in 1:maxdepth
        idx = findall(G.depths .== lnow)
        verbose && @printf("Training level %d/%d... #nodes = %d, ", lnow, maxdepth, length(idx))
        if lnow == 1
            @assert length(idx) == 1 "Only 1 root node (level = 1) allowed."
            @assert G.depths[1] == 1 "Root node depth must be in the 1st place."
        end
        if !isempty(idx)
            X = get_x(G, G.levels[idx, :], G.indices[idx, :])
            fv = Float64[f(X[k, :]) for k in 1:length(idx)]
            al = Vector{Float64}(undef, length(idx)); idx0 = Int64.(idx .- 1)
            # alpha = f - G_{depth < lnow}, stored as the coefficients of this depth on the device too
            check(ccall((:asg_surplus_nodes, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}),
                h, length(idx), idx0, fv, lnow - 1, 1, al))
            G.fvals[idx] = fv; G.coefs[idx] = al          # in-place: mirror stays valid
        end
        verbose && @printf("Elapsed: %.2f seconds.\n", time() - tic)
    end
    verbose && @printf("done in %.2f seconds.\n", time() - tic)
    return nothing
end

"""
    children(G, depth) -> (levels, indices, X)

Unique valid children of all nodes of `depth` (src/train.jl:337-398), generated on the device, with coordinates.
"""
function children(G::SparseGridInterpolant{D}, depth::Int) where {D}
    h = ensure!(G)
    np_, nc, nu = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:asg_children, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}, Ref{Int64}), h, depth, np_, nc, nu))
    n = nu[]
    L = Matrix{Int}(undef, n, D); I = Matrix{Int}(undef, n, D); X = Matrix{Float64}(undef, n, D)
    n > 0 && GC.@preserve L I X check(ccall((:asg_children_fetch, LIB), Int32,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Ptr{Float64}, Int64, Int64), h, n, L, I, 1, n, X, 1, n))
    return L, I, X
end

function adapt!(G::SparseGridInterpolant{D}, f2fit::Function, maxdepth::Int; verbose::Bool = true, tol::Float64 = 1e-3,
                toltype::Symbol = :absolute, parallel::Bool = false) where {D}
    @assert length(G) > 0 "The grid structure in G is empty."
    lnow = maximum(G.depths)
    @assert lnow <= maxdepth "maxdepth must be >= current depth."
    @assert lnow > 0 "Ill-defined grid structure: maxdepth < 1 found."
    toltype in (:absolute, :relative) || throw(ArgumentError("toltype must be :absolute or :relative."))
    big(a, fx) = toltype == :relative ? (abs(fx) <= tol ? abs(a) > tol : abs(a) / abs(fx) > tol) : abs(a) > tol
    verbose && println("Adaptive sparse grid adaption starts\n", "# dimensions: $D, # current total nodes: $(length(G)), ",
                       toltype, " tolerance: ", tol)
    tic = time()
    if lnow >= maxdepth
        verbose && @printf("Current depth %d >= max depth %d, stop adaption.\n", lnow, maxdepth)
        return nothing
    end
    while lnow < maxdepth
        parents = findall(G.depths .== lnow)
        verbose && @printf("level %d/%d, #parents: %d, ", lnow, maxdepth, length(parents))
        # children, validity test and removal of repeats on the device (asg_children): every unique child is
        # evaluated once; rows in the reference's visiting order (for j, for parent, left then right)
        CL, CI, X = children(G, lnow)
        new = Dict{NTuple{2,Vector{Int}},NTuple{2,Float64}}(); order = NTuple{2,Vector{Int}}[]
        if size(CL, 1) > 0
            fv = Float64[f2fit(X[k, :]) for k in 1:size(X, 1)]
            al = surplus(G, X, fv)                                    # ONE kernel call for the whole level
            for k in eachindex(fv)
                key = (CL[k, :], CI[k, :])
                if big(al[k], fv[k]) && !haskey(new, key); new[key] = (fv[k], al[k]); push!(order, key); end
            end
        end
        if isempty(new)
            verbose && @printf("\nConverged: no new nodes found. Elapsed: %.2f seconds\n", time() - tic)
            break
        end
        n = length(order)
        nl = permutedims(reduce(hcat, first.(order))); ni = permutedims(reduce(hcat, last.(order)))
        nf = [new[k][1] for k in order]; na = [new[k][2] for k in order]
        GC.@preserve nl ni na check(ccall((:asg_grid_append, LIB), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Ptr{Float64}), G.handle, n, nl, ni, 1, n, na))
        setfield!(G, :levels, vcat(G.levels, nl)); setfield!(G, :indices, vcat(G.indices, ni))
        setfield!(G, :depths, vcat(G.depths, [nodedepth(nl[k, :]) for k in 1:n]))
        setfield!(G, :fvals, vcat(G.fvals, nf)); setfield!(G, :coefs, vcat(G.coefs, na))
        verbose && @printf("#added/#candidates/#total: %d/%d/%d, Elapsed: %.2f seconds.\n", n, length(parents) * 2 * D,
                           length(G), time() - tic)
        lnow += 1
        if lnow == maxdepth
            verbose && @printf("Reached the maximum depth %d. Elapsed: %.2f seconds.\n", maxdepth, time() - tic)
            break
        end
    end
    return nothing
end

end # module
