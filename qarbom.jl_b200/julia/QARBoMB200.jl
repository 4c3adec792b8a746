# QARBoMB200.jl -- drop-in Julia shim for the classical RBM training hot path of QARBoM.jl.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary, so this file is
# shipped unverified.  What IS verified (pytest, on a B200) is the C ABI it binds
# (include/qarbom_b200.h) through the Python ctypes mirror in qarbom.jl_b200/*.py, which follows
# the same call sequence.  See INTEGRATION.md.
#
# It keeps the reference's public surface for this path: the model types and their fields
# (src/rbm.jl:2-39), `train!(rbm, x_train, CD|PCD; kwargs...)` and the classifier variants
# (src/training/train_cd.jl:45-59,121-138; src/training/train_pcd.jl:133-148,258-276),
# `reconstruct` (src/rbm.jl:220-224), the MeanSquaredError metric and `_diverged`
# (src/evaluation.jl:16-20,132-142), the epoch log and CSV schema (src/utils.jl:28-43,
# src/training/train_cd.jl:113-115).  The QUBO / quantum-sampler path is not touched: keep using
# QARBoM.jl for `QSampling`.
module QARBoMB200

using Printf

export RBM, GRBM, RBMClassifier, GRBMClassifier, CD, PCD, FastPCD, MeanSquaredError, FreeEnergy

const LIB = get(ENV, "QARBOM_B200_LIB", joinpath(@__DIR__, "..", "lib", "libqarbom_b200.so"))

abstract type TrainingMethod end
abstract type PCD <: TrainingMethod end
abstract type CD <: TrainingMethod end
abstract type FastPCD <: TrainingMethod end
abstract type EvaluationMethod end
abstract type MeanSquaredError <: EvaluationMethod end
abstract type FreeEnergy <: EvaluationMethod end   # new metric
abstract type AbstractRBM end

mutable struct RBM <: AbstractRBM
    W::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; n_visible::Int; n_hidden::Int
end
mutable struct GRBM <: AbstractRBM
    W::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; n_visible::Int; n_hidden::Int
end
mutable struct RBMClassifier <: AbstractRBM
    W::Matrix{Float64}; U::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; c::Vector{Float64}
    n_visible::Int; n_hidden::Int; n_classifiers::Int
end
mutable struct GRBMClassifier <: AbstractRBM
    W::Matrix{Float64}; U::Matrix{Float64}; a::Vector{Float64}; b::Vector{Float64}; c::Vector{Float64}
    n_visible::Int; n_hidden::Int; n_classifiers::Int
end
const RBMClassifiers = Union{RBMClassifier, GRBMClassifier}

RBM(v::Int, h::Int) = RBM(randn(v, h), zeros(v), zeros(h), v, h)
RBM(v::Int, h::Int, W::Matrix{Float64}) = RBM(copy(W), zeros(v), zeros(h), v, h)
GRBM(v::Int, h::Int) = GRBM(randn(v, h), zeros(v), zeros(h), v, h)
GRBM(v::Int, h::Int, W::Matrix{Float64}) = GRBM(copy(W), zeros(v), zeros(h), v, h)
RBMClassifier(v::Int, h::Int, l::Int) = RBMClassifier(randn(v, h), randn(l, h), zeros(v), zeros(h), zeros(l), v, h, l)
RBMClassifier(v::Int, h::Int, l::Int, W::Matrix{Float64}, U::Matrix{Float64}) =
    RBMClassifier(copy(W), copy(U), zeros(v), zeros(h), zeros(l), v, h, l)
GRBMClassifier(v::Int, h::Int, l::Int) = GRBMClassifier(randn(v, h), randn(l, h), zeros(v), zeros(h), zeros(l), v, h, l)
GRBMClassifier(v::Int, h::Int, l::Int, W::Matrix{Float64}, U::Matrix{Float64}) =
    GRBMClassifier(copy(W), copy(U), ones(v) * 1e-5, ones(h) * 1e-5, ones(l) * 1e-5, v, h, l)

n_labels(::AbstractRBM) = 0
n_labels(r::RBMClassifiers) = r.n_classifiers
is_gaussian(::AbstractRBM) = false
is_gaussian(::Union{GRBM, GRBMClassifier}) = true

# ------------------------------------------------------------------------------ C ABI
struct QbConfig
    n_visible::Int64; n_hidden::Int64; n_classifiers::Int64
    visible_kind::Int32; precision::Int32; max_batch::Int64
    device::Int32; rank::Int32; world::Int32; reserved::Int32
end

const QB_DT_F64 = Int32(0); const QB_DT_I64 = Int32(1)

function check(rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:qb_last_error, LIB), Cstring, ()))
    error("qarbom_b200 error $rc: $msg")
end

mutable struct Device
    h::Ptr{Cvoid}
    function Device(rbm::AbstractRBM, max_batch::Int; precision::Symbol = :bf16, device::Int = 0, seed::UInt64 = UInt64(0))
        cfg = Ref(QbConfig(rbm.n_visible, rbm.n_hidden, n_labels(rbm), is_gaussian(rbm) ? 1 : 0,
                           precision == :fp32 ? 0 : 1, max_batch, device, 0, 1, 0))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qb_create, LIB), Int32, (Ref{Ptr{Cvoid}}, Ref{QbConfig}), out, cfg))
        d = new(out[])
        finalizer(x -> (x.h != C_NULL && ccall((:qb_destroy, LIB), Int32, (Ptr{Cvoid},), x.h); x.h = C_NULL), d)
        push_params!(d, rbm)
        check(ccall((:qb_seed, LIB), Int32, (Ptr{Cvoid}, UInt64), d.h, seed))
        return d
    end
end

_U(r) = n_labels(r) > 0 ? pointer(r.U) : Ptr{Float64}(C_NULL)
_c(r) = n_labels(r) > 0 ? pointer(r.c) : Ptr{Float64}(C_NULL)

function push_params!(d::Device, r::AbstractRBM)   # Matrix{Float64} is already column-major V x H
    GC.@preserve r check(ccall((:qb_set_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        d.h, pointer(r.W), pointer(r.a), pointer(r.b), _U(r), _c(r)))
end
function pull_params!(d::Device, r::AbstractRBM)
    GC.@preserve r check(ccall((:qb_get_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        d.h, pointer(r.W), pointer(r.a), pointer(r.b), _U(r), _c(r)))
end

# Vector{Vector{T}} -> dense V x N matrix (column j = sample j), i.e. N rows of V contiguous elements
_dense(x::Vector{Vector{Int}}) = (reduce(hcat, x), QB_DT_I64)
_dense(x::Vector{Vector{Float64}}) = (reduce(hcat, x), QB_DT_F64)

function set_data!(d::Device, x, y = nothing)
    X, xdt = _dense(x)
    if y === nothing
        GC.@preserve X check(ccall((:qb_set_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64),
                                   d.h, pointer(X), xdt, C_NULL, 0, size(X, 2)))
    else
        Y, ydt = _dense(y)
        GC.@preserve X Y check(ccall((:qb_set_data, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64),
                                     d.h, pointer(X), xdt, pointer(Y), ydt, size(X, 2)))
    end
end

function eval_mse(d::Device, x = nothing)
    out = Ref{Float64}(0.0)
    if x === nothing
        check(ccall((:qb_eval_mse, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{Float64}),
                    d.h, C_NULL, 0, 0, C_NULL, out))
    else
        X, xdt = _dense(x)
        GC.@preserve X check(ccall((:qb_eval_mse, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ref{Float64}),
                                   d.h, pointer(X), xdt, size(X, 2), C_NULL, out))
    end
    return out[]
end

function epoch!(d::Device, ::Type{PCD}, B, k, lr, llr)
    t = zeros(3)
    check(ccall((:qb_pcd_epoch, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Ptr{Float64}), d.h, B, k, lr, llr, t))
    return t[1], t[2], t[3]
end
function epoch!(d::Device, ::Type{CD}, B, k, lr, llr)
    t = zeros(3)
    check(ccall((:qb_cd_epoch, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Ptr{Float64}), d.h, B, k, lr, llr, t))
    return t[1], t[2], t[3]
end

"gibbs_sample_hidden / gibbs_sample_visible for a batch of rows (columns of the V x n / H x n matrices); Philox stream"
function sample_hidden(d::Device, rbm::AbstractRBM, v::Matrix{Float64})
    out = zeros(rbm.n_hidden, size(v, 2))
    GC.@preserve v out check(ccall((:qb_sample_hidden, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                   d.h, pointer(v), C_NULL, size(v, 2), C_NULL, pointer(out)))
    return out
end
function sample_visible(d::Device, rbm::AbstractRBM, h::Matrix{Float64})
    out = zeros(rbm.n_visible, size(h, 2))
    GC.@preserve h out check(ccall((:qb_sample_visible, LIB), Int32,
                                   (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                   d.h, pointer(h), size(h, 2), C_NULL, C_NULL, pointer(out), C_NULL))
    return out
end

"fast_persistent_contrastive_divergence! for one epoch (src/training/train_fast_pcd.jl:1-57, classifiers :59-127)"
function fast_pcd_epoch!(d::Device, B, k, lr, fast_lr, llr = 0.0, fast_llr = 0.0)
    t = zeros(3)
    check(ccall((:qb_fast_pcd_epoch, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Float64, Float64, Float64, Float64, Ptr{Float64}),
                d.h, B, k, lr, fast_lr, llr, fast_llr, t))
    return t[1], t[2], t[3]
end

"One mini-batch of persistent_qubo! (src/training/train_quantum_pcd.jl:10-37): the annealer's (v~, h~) come from Julia"
function qsampling_step!(d::Device, batch_index::Int, B::Int, v_model::Vector{Float64}, h_model::Vector{Float64}, lr::Float64)
    GC.@preserve v_model h_model check(ccall((:qb_init_chains, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                               d.h, 1, pointer(v_model), pointer(h_model), C_NULL))
    check(ccall((:qb_pcd_step, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int64, Float64, Float64), d.h, batch_index, B, 0, lr, 0.0))
end

snapshot!(d::Device) = check(ccall((:qb_snapshot_params, LIB), Int32, (Ptr{Cvoid},), d.h))   # best_rbm = copy_rbm(rbm), in HBM
restore!(d::Device) = check(ccall((:qb_restore_params, LIB), Int32, (Ptr{Cvoid},), d.h))     # copy_rbm!(best_rbm, rbm)

# ------------------------------------------------------------------------------ reference plumbing kept in Julia
function _diverged(h::Vector{Float64})           # src/evaluation.jl:132-142
    length(h) == 1 && return false
    return h[end] > h[end-1] || h[end] > minimum(h)
end

function _log_epoch(epoch, t_sample, t_gibbs, t_update, total_t)   # src/utils.jl:28-43
    println("|------------------------------------------------------------------|")
    println("| Epoch | Time (Sample) | Time (Gibbs) | Time (Update) | Total     |")
    println("|------------------------------------------------------------------|")
    println(@sprintf("| %5d | %13.4f | %12.4f | %13.4f | %9.4f |", epoch, t_sample, t_gibbs, t_update, total_t))
    println("|------------------------------------------------------------------|")
end

copy_params(r::AbstractRBM) = n_labels(r) > 0 ? (copy(r.W), copy(r.a), copy(r.b), copy(r.U), copy(r.c)) :
                                                (copy(r.W), copy(r.a), copy(r.b))
function restore_params!(r::AbstractRBM, p)
    r.W .= p[1]; r.a .= p[2]; r.b .= p[3]
    if n_labels(r) > 0
        r.U .= p[4]; r.c .= p[5]
    end
end

"""
    train!(rbm, x_train, CD | PCD | FastPCD; n_epochs, learning_rate, [batch_size], [gibbs_steps], ...)
    train!(rbm::RBMClassifiers, x_train, label_train, CD | PCD | FastPCD; ..., label_learning_rate, ...)

Same keyword arguments as QARBoM.train! for CD / PCD / FastPCD (`fast_learning_rate`, `fast_label_learning_rate`); the per-epoch Gibbs chains, gradient,
update and MSE evaluation run on the B200.  New keywords (`batch_size` for CD, `momentum`,
`weight_decay`, `precision`, `seed`) default to the reference's behaviour.
"""
function train!(rbm::AbstractRBM, x_train, method::Type{<:TrainingMethod}; label_train = nothing,
                n_epochs::Int, learning_rate::Vector{Float64}, label_learning_rate::Vector{Float64} = learning_rate,
                fast_learning_rate::Float64 = 0.1, fast_label_learning_rate::Float64 = 0.1,
                gibbs_steps::Int = method === CD ? 3 : 1, batch_size::Int = method === CD ? 1 : error("batch_size required"),
                metrics = [MeanSquaredError], early_stopping::Bool = false, store_best_rbm::Bool = true,
                patience::Int = 10, stopping_metric = MeanSquaredError, x_test_dataset = nothing,
                file_path = method === FastPCD ? "fast_pcd_metrics.csv" : (method === PCD ? "pcd_metrics.csv" : "cd_metrics.csv"),
                momentum::Float64 = 0.0, weight_decay::Float64 = 0.0, precision::Symbol = :bf16, seed::UInt64 = UInt64(0))
    N = length(x_train)
    if N % batch_size > 0   # _set_mini_batches, src/utils.jl:13-26
        @warn "The last batch size is not equal to the other batches. Will dismiss $(N % batch_size) samples."
    end
    d = Device(rbm, max(batch_size, min(N, 4096)); precision = precision, seed = seed)
    check(ccall((:qb_set_optimizer, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), d.h, momentum, weight_decay))
    set_data!(d, x_train, label_train)
    best = copy_params(rbm)
    initial_patience = patience
    mse = Float64[eval_mse(d, x_test_dataset)]          # initial_evaluation
    if method !== CD
        println("Setting mini-batches")
        check(ccall((:qb_init_chains, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    d.h, batch_size, C_NULL, C_NULL, C_NULL))   # _init_fantasy_data(rbm, batch_size)
        println("Starting training")
    end
    tot = (0.0, 0.0, 0.0)
    hist = Float64[]
    for epoch in 1:n_epochs
        t = method === FastPCD ?
            fast_pcd_epoch!(d, batch_size, gibbs_steps, learning_rate[epoch], fast_learning_rate, label_learning_rate[epoch],
                            fast_label_learning_rate) :
            epoch!(d, method, batch_size, gibbs_steps, learning_rate[epoch], label_learning_rate[epoch])
        tot = tot .+ t
        push!(hist, eval_mse(d, x_test_dataset))
        if _diverged(hist)
            if early_stopping
                if patience == 0
                    println("Early stopping at epoch $epoch")
                    break
                end
                patience -= 1
            end
        else
            patience = initial_patience
            if store_best_rbm
                pull_params!(d, rbm); best = copy_params(rbm)
            end
        end
        _log_epoch(epoch, t[1], t[2], t[3], sum(tot))
        println("mse: $(hist[end])")
    end
    store_best_rbm ? restore_params!(rbm, best) : pull_params!(d, rbm)
    open(file_path, "w") do io           # CSV.write(file_path, DataFrame(metrics_dict)) schema: one column per metric
        println(io, "mse")
        foreach(v -> println(io, v), vcat(mse, hist))
    end
    finalize(d)
    return
end

train!(rbm::RBMClassifiers, x_train, label_train, method::Type{<:TrainingMethod}; kwargs...) =
    train!(rbm, x_train, method; label_train = label_train, kwargs...)

"reconstruct(rbm, v), src/rbm.jl:220-224"
function reconstruct(rbm::AbstractRBM, v::Vector{<:Number})
    d = Device(rbm, 1; precision = :fp32)
    vin = Float64.(v); out = similar(vin)
    GC.@preserve vin out check(ccall((:qb_reconstruct, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                     d.h, pointer(vin), 1, C_NULL, pointer(out)))
    finalize(d)
    return out
end

end # module
