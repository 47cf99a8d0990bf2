# IMCB200.jl — Julia shim that routes the reference's transport-step call sites to libimc_b200.so.
#
# Drop it next to the reference's src/ and replace the six `include("imc_*.jl")` lines of the stage
# modules in src/MixedPrecisionIMC.jl (:13-18) by `include("IMCB200.jl")`: the module and function names,
# argument lists and mutation behaviour below are the reference's (imc_update.jl:12, imc_sourcing.jl:12,
# imc_transport.jl:13/212/483/786, imc_clean.jl:6, imc_tally.jl:11, imc_energycheck.jl:10), so `main`
# (MixedPrecisionIMC.jl:59-179), the deck parser, the mesh generator, time stepping and output are unchanged.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not available in the build environment.  The Python twin of
# this file (mixedprecisionimc.jl_b200/driver.py + lib.py) is what the tests drive; the struct layouts and
# argument orders below mirror include/imc.h one to one.
module IMCB200

const libimc = get(ENV, "IMC_B200_LIB", "libimc_b200.so")
const IMC_MAX_SCALES = 16

struct ImcConfig            # imc_config (include/imc.h)
    struct_size::Int32; precision::Int32; geometry::Int32; nx::Int32; ny::Int32
    bc::NTuple{4,Int32}
    linearized::Int32; pairwise::Int32; randomwalk::Int32; marshak_quirk::Int32; n_scales::Int32
    energyscales::NTuple{IMC_MAX_SCALES,Float64}
    distancescale::Float64; phys_c::Float64; phys_a::Float64; alpha::Float64
    seed::Int64; n_max::Int64
    device::Int32; rank::Int32; world::Int32
    rng_mode::Int32; tally_mode::Int32; track_mode::Int32
    exact_record_budget::Int64
end
mutable struct SourceStats; totalenergy::Float64; emitted_sum::Float64; n_source::Int64; n_new_global::Int64; n_new_local::Int64; n_particles::Int64; SourceStats() = new(0, 0, 0, 0, 0, 0); end
mutable struct TransportStats; lostenergy::Float64; segments::UInt64; segments_total::UInt64; histories::Int64; n_census::Int64; n_absorbed::Int64; n_escaped::Int64; n_rw::Int64; n_errors::Int64; variant::Int32; tally_mode::Int32; kernel_ms::Float32; TransportStats() = new(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0); end
mutable struct TallyStats; totalenergydep::Float64; energy_increase::Float64; max_temp::Float64; total_energy_density::Float64; TallyStats() = new(0, 0, 0, 0); end
mutable struct EnergyStats; radenergy::Float64; radenergy_change::Float64; lostenergy::Float64; energy_error::Float64; EnergyStats() = new(0, 0, 0, 0); end

const ENGINES = IdDict{Any,Ptr{Cvoid}}()   # mesh => imc_handle
const STEP = Ref(0)                        # time-step ordinal (Philox counter word)
const PENDING_RW = Ref{Any}(nothing)       # (aVals, prVals, ptVals) of Transport.randomwalk_table, until the engine exists

check(h, rc) = rc == 0 || error("imc error $rc: " * unsafe_string(ccall((:imc_last_error, libimc), Cstring, (Ptr{Cvoid},), h)))
f64(a) = Float64.(vec(collect(a)))         # column-major linear order, as the ABI expects
bcid(s) = uppercase(string(s)) == "REFLECT" ? Int32(0) : uppercase(string(s)) == "VACUUM" ? Int32(1) : error("BC must be REFLECT or VACUUM")
precid(T) = T === Float16 ? Int32(0) : T === Float32 ? Int32(1) : Int32(2)

"""Create the engine for this deck and upload the mesh (first call of Update.update does it lazily)."""
function attach!(inputs, mesh, simvars)
    geom = simvars.geometry == "1D" ? 1 : 2
    nx, ny = geom == 1 ? (Int(mesh.Ncells), 1) : Tuple(Int.(mesh.Ncells))
    scales = sort(Float64.(collect(mesh.energyscales)), rev=true)
    bc = geom == 1 ? (bcid(simvars.BC[1]), bcid(simvars.BC[2]), Int32(1), Int32(1)) :
                     (bcid(simvars.BC[1]), bcid(simvars.BC[2]), bcid(simvars.BC[3]), bcid(simvars.BC[4]))
    C = parentmodule(@__MODULE__).Constants     # the module that included this file (MixedPrecisionIMC)
    cfg = ImcConfig(sizeof(ImcConfig), precid(simvars.precision), geom, nx, ny, bc,
        inputs["LINEARIZED"] == "TRUE", simvars.pairwise == "TRUE",
        geom == 1 && uppercase(string(get(inputs, "RANDOMWALK", "FALSE"))) == "TRUE",
        uppercase(inputs["NAME"]) == "MARSHAK WAVE", length(scales),
        ntuple(i -> i <= length(scales) ? scales[i] : 1.0, IMC_MAX_SCALES),
        Float64(mesh.distancescale), Float64(C.phys_c), Float64(C.phys_a), Float64(C.alpha),
        parse(Int, inputs["SEED"]), simvars.n_max, 0, 0, 1, 0, 0, 0, 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:imc_create, libimc), Cint, (Ref{ImcConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error("imc_create: " * unsafe_string(ccall((:imc_last_error, libimc), Cstring, (Ptr{Cvoid},), C_NULL)))
    sel(a, k) = geom == 1 ? a[:, k] : a[:, :, k]
    ts = mesh.temp_surf
    args = geom == 1 ?
        (f64(mesh.dx), Float64[], f64(sel(mesh.sigma_a, 2)), f64(sel(mesh.sigma_a, 3)), f64(sel(mesh.sigma_s, 2)), f64(sel(mesh.sigma_s, 3)),
         f64(mesh.sigma[:, 1]), f64(mesh.bee), f64(mesh.radsource), f64(mesh.temp), Float64[], Float64[], [Float64(ts[1])], [Float64(ts[2])]) :
        (f64(mesh.dx), f64(mesh.dy), f64(sel(mesh.sigma_a, 2)), f64(sel(mesh.sigma_a, 3)), f64(sel(mesh.sigma_s, 2)), f64(sel(mesh.sigma_s, 3)),
         f64(mesh.sigma[:, :, 1]), f64(mesh.bee), f64(mesh.radsource), f64(mesh.temp), f64(ts[1]), f64(ts[2]), f64(ts[3]), f64(ts[4]))
    # ccall takes a literal tuple of argument types and no splatting, hence the fourteen names
    (a_dx, a_dy, a_sac, a_sap, a_ssc, a_ssp, a_sig, a_bee, a_rad, a_temp, a_tsb, a_tst, a_tsl, a_tsr) = args
    PF = Ptr{Float64}
    check(h[], ccall((:imc_set_mesh, libimc), Cint, (Ptr{Cvoid}, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF, PF),
                     h[], a_dx, a_dy, a_sac, a_sap, a_ssc, a_ssp, a_sig, a_bee, a_rad, a_temp, a_tsb, a_tst, a_tsl, a_tsr))
    if PENDING_RW[] !== nothing                # random-walk tables: built by the engine, copied into the host's arrays
        aVals, prVals, ptVals = PENDING_RW[]
        a = Vector{Float64}(undef, length(aVals)); pr = similar(a); pt = similar(a)
        check(h[], ccall((:imc_rw_table, libimc), Cint, (Ptr{Cvoid}, Float64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                         h[], Float64(first(aVals)), Float64(last(aVals)), length(aVals), a, pr, pt))
        prVals .= pr; ptVals .= pt
        PENDING_RW[] = nothing
    end
    ENGINES[mesh] = h[]
    STEP[] = 0
    return h[]
end
engine(inputs, mesh, simvars) = get!(() -> attach!(inputs, mesh, simvars), ENGINES, mesh)
engine(mesh) = ENGINES[mesh]

function pull!(mesh, field::Symbol, id::Integer)   # imc_get_field_native straight into the reference's Array{T}
    dst = getfield(mesh, field)
    h = engine(mesh)
    es = ccall((:imc_field_elsize, libimc), Int32, (Ptr{Cvoid}, Int32), h, id)
    if es != sizeof(eltype(dst))                     # mesh.temp turns Float64 after the first LINEARIZED tally (imc_tally.jl:72)
        dst = Array{es == 8 ? Float64 : (es == 4 ? Float32 : Float16)}(undef, size(dst))
        setfield!(mesh, field, dst)
    end
    GC.@preserve dst check(h, ccall((:imc_get_field_native, libimc), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64), h, id, pointer(dst), sizeof(dst)))
    dst
end

"""Stand-in for the reference's `particles` vector: the engine owns the population."""
struct ParticleHandle <: AbstractVector{Vector{Float64}}; mesh; end
Base.size(p::ParticleHandle) = (Int(ccall((:imc_num_particles, libimc), Int64, (Ptr{Cvoid},), engine(p.mesh))),)

module Update
    import ..IMCB200: engine, check, libimc
    function update(inputs, mesh, simvars)                                   # imc_update.jl:12
        h = engine(inputs, mesh, simvars)
        check(h, ccall((:imc_update, libimc), Cint, (Ptr{Cvoid}, Float64), h, Float64(simvars.dt)))
    end
end

module Sourcing
    import ..IMCB200: engine, check, libimc, SourceStats, STEP
    function sourcing(mesh, simvars, particles)                               # imc_sourcing.jl:12
        st = SourceStats()
        check(engine(mesh), ccall((:imc_source, libimc), Cint, (Ptr{Cvoid}, Float64, Int64, Float64, Int64, Int64, Ref{SourceStats}),
              engine(mesh), Float64(simvars.dt), simvars.n_input, Float64(simvars.cellmin), STEP[], -1, st))
        mesh.totalenergy = simvars.precision(st.totalenergy)
        print("Total intial time-step energy ", st.emitted_sum + sum(mesh.radenergydens), "\n")
        print("The number of particles after sourcing is ", st.n_particles, "\n")
    end
end

"""`Sourcing.sample_planck` (imc_sourcing.jl:372-399; every call site in the reference is commented out): n Planck-spectrum
frequencies drawn by the engine in the deck precision."""
function sample_planck(mesh, n::Integer)
    out = Vector{Float64}(undef, n)
    check(engine(mesh), ccall((:imc_sample_planck, libimc), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}), engine(mesh), n, STEP[], out))
    return eltype(mesh.matenergydens).(out)
end

module Transport
    import ..IMCB200: engine, check, libimc, TransportStats, STEP, PENDING_RW
    function run(mesh, simvars)
        st = TransportStats()
        check(engine(mesh), ccall((:imc_transport, libimc), Cint, (Ptr{Cvoid}, Float64, Int64, Ref{TransportStats}), engine(mesh), Float64(simvars.dt), STEP[], st))
        simvars.iterations = Int(st.segments_total)
        mesh.lostenergy = simvars.precision(st.lostenergy)
        print("There were ", simvars.iterations, " total iterations this time-step. \n")
    end
    MC(mesh, simvars, particles) = run(mesh, simvars)                         # imc_transport.jl:13
    MC_RW(mesh, simvars, rwvars, particles) = run(mesh, simvars)              # imc_transport.jl:212
    MC2D(mesh, simvars, particles) = run(mesh, simvars)                       # imc_transport.jl:483
    # imc_transport.jl:786.  `main` calls this before the first Update.update (MixedPrecisionIMC.jl:132), i.e. before the
    # engine exists: the request is parked and `attach!` fills prVals / ptVals in place (RWVars holds these same arrays).
    function randomwalk_table(aVals, prVals, ptVals, simvars)
        PENDING_RW[] = (aVals, prVals, ptVals)
        return prVals, ptVals
    end
end

module Clean
    import ..IMCB200: engine, check, libimc, ParticleHandle
    function clean(particles::ParticleHandle)                                 # imc_clean.jl:6
        n = Ref{Int64}(0)
        check(engine(particles.mesh), ccall((:imc_clean, libimc), Cint, (Ptr{Cvoid}, Ref{Int64}), engine(particles.mesh), n))
    end
end

module Tally
    import ..IMCB200: engine, check, libimc, TallyStats, pull!
    function tally(inputs, mesh, simvars, particles)                          # imc_tally.jl:11
        st = TallyStats()
        check(engine(mesh), ccall((:imc_tally, libimc), Cint, (Ptr{Cvoid}, Float64, Float64, Ref{TallyStats}), engine(mesh), Float64(simvars.t), Float64(simvars.dt), st))
        mesh.totalenergydep = simvars.precision(st.totalenergydep)
        pull!(mesh, :temp, 0); pull!(mesh, :matenergydens, 8); pull!(mesh, :radenergydens, 9); pull!(mesh, :energydep, 6)
        print("Energy increase: ", st.energy_increase, "\n")
        print("Maximum mesh temperature is ", st.max_temp, "\n")
        print("Final total energy density ", st.total_energy_density, "\n")
        nrg_inc = similar(mesh.matenergydens)                                 # imc_tally.jl:58
        GC.@preserve nrg_inc check(engine(mesh), ccall((:imc_get_field_native, libimc), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64), engine(mesh), 10, pointer(nrg_inc), sizeof(nrg_inc)))
        push!(mesh.energyincrease_saved, nrg_inc)
        push!(mesh.temp_saved, copy(mesh.temp)); push!(mesh.matenergy_saved, copy(mesh.matenergydens)); push!(mesh.radenergy_saved, copy(mesh.radenergydens))
    end
end

"""Deferred variant of the per-step pulls: `IMCB200.history_enable(mesh, nsteps)` once, then `Tally.tally` may skip
`pull!`/`push!` and `IMCB200.fetch_history!(mesh)` appends every recorded step to the reference's `*_saved` lists at
the end of the run (imc_history_* in include/imc.h)."""
history_enable(mesh, nsteps::Integer) = check(engine(mesh), ccall((:imc_history_enable, libimc), Cint, (Ptr{Cvoid}, Int64), engine(mesh), nsteps))
function fetch_history!(mesh)
    h = engine(mesh); n = Ref{Int64}(0); dropped = Ref{Int64}(0)
    check(h, ccall((:imc_history_count, libimc), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), h, n, dropped))
    T = eltype(mesh.matenergydens); shape = size(mesh.matenergydens)
    for (id, lst, E) in ((0, mesh.temp_saved, Float64), (8, mesh.matenergy_saved, T), (9, mesh.radenergy_saved, T), (10, mesh.energyincrease_saved, T))
        buf = Array{E}(undef, prod(shape), n[])
        GC.@preserve buf check(h, ccall((:imc_history_get, libimc), Cint, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Cvoid}, Int64), h, id, 0, n[], pointer(buf), sizeof(buf)))
        for k in 1:n[]; push!(lst, reshape(buf[:, k], shape)); end
    end
    check(h, ccall((:imc_history_clear, libimc), Cint, (Ptr{Cvoid},), h))
    return n[]
end

"""Restart point inside the library (imc_checkpoint, include/imc.h): `checkpoint!(mesh, :save)` copies the engine's whole
mutable state — what `deepcopy(mesh), deepcopy(particles)` would keep in the reference — in device memory, `:restore` makes it
current again, `:drop` frees it.  The host restores its own `simvars.t`, `simvars.dt` and `STEP[]`."""
checkpoint!(mesh, op::Symbol) = check(engine(mesh), ccall((:imc_checkpoint, libimc), Cint, (Ptr{Cvoid}, Int32), engine(mesh),
                                                         Int32(op === :save ? 0 : op === :restore ? 1 : 2)))

module EnergyCheck
    import ..IMCB200: engine, check, libimc, EnergyStats, STEP
    function energychecker(inputs, mesh, simvars, particles)                  # imc_energycheck.jl:10
        st = EnergyStats()
        check(engine(mesh), ccall((:imc_energycheck, libimc), Cint, (Ptr{Cvoid}, Ref{EnergyStats}), engine(mesh), st))
        print("Total energy: ", mesh.totalenergy, " Total energy deposition: ", mesh.totalenergydep, " Radiation energy change: ", st.radenergy_change, " Lost energy: ", st.lostenergy, "\n")
        print("The energy conservation error is: ", st.energy_error, " \n")
        mesh.radenergyold = simvars.precision(st.radenergy); mesh.lostenergy = simvars.precision(0.0)
        STEP[] += 1                                                           # last stage of the step
    end
end

end # module
