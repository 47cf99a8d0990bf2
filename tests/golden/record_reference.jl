#!/usr/bin/env julia
# record_reference.jl — pins the oracle against the REAL reference, on a machine that has Julia.
#
#   julia tests/golden/record_reference.jl <checkout of simonbutson/MixedPrecisionIMC.jl> <deck.txt> <out dir> [steps]
#
# NOT EXECUTED IN THIS REPOSITORY: Julia is not installed in its build environment (SURVEY.md §8c), so parity against the
# real reference is unpinned (DESIGN.md §5).  This script is the missing half: it runs the reference's own `main` for a few
# time steps on a small deck and records, per stage and step, every `rand` / `randexp` draw (attributed to the particle
# that consumed it) and the state the stage left behind, in the plain-binary format that
# tests/golden/recorded_run.py::replay feeds to the oracle and to the CUDA engine (replay mode of BASELINE.json:
# events / cell indices exact, energies and positions within 4 ulp).  The Python side of the format is exercised by
# tests/test_golden.py::test_recorded_run_format with a fixture written by the second restatement (oracle/imc_refpy.py).
#
# How it records without touching the algorithms: the reference's sources are copied to a temporary directory and three
# of them are patched TEXTUALLY — `rand(` / `randexp(` become calls of the recording wrappers below (which call the real
# Random functions and return their values unchanged), the particle loops of imc_transport.jl announce the particle they
# work on, and `main` gets hooks between its stages, a step limit, and loses its plotting call.  Nothing else changes.
# Put the resulting directory under tests/golden/julia/<name>/ and the test-suite picks it up.

module REC
using Random
const mode = Ref(:off)                       # :source | :transport
const cur = Ref(0)                           # transport: index of the particle being tracked
const base = Ref(0)                          # sourcing: length(particles) before the call
const uni = Dict{Int,Vector{Float64}}()      # slot => draws in call order
const ex = Dict{Int,Vector{Float64}}()
const outdir = Ref("")
const nsteps = Ref(3)
const step = Ref(0)
const scal = Float64[]

more() = step[] < nsteps[]
note!(d, slot, v) = (push!(get!(d, slot, Float64[]), Float64(v)); v)
srand(particles, T) = note!(uni, length(particles) - base[] + 1, Random.rand(T))       # imc_sourcing.jl: slot = ordinal of the new particle
trand(T) = note!(uni, cur[], Random.rand(T))                                           # imc_transport.jl
trandexp(T) = note!(ex, cur[], Random.randexp(T))
trandexp() = note!(ex, cur[], Random.randexp())

path(name) = joinpath(outdir[], string(step[], "_", name))
put(name, a) = open(io -> write(io, Float64.(vec(collect(a)))), path(name * ".f64"), "w")
function dump_ragged(name, d, n)
    open(io -> write(io, Int64[length(get(d, i, Float64[])) for i in 1:n]), path(name * ".len"), "w")
    open(io -> foreach(i -> write(io, get(d, i, Float64[])), 1:n), path(name * ".f64"), "w")
end
function dump_particles(name, particles)      # row-major [n, nslots], every slot as Float64
    open(path(name * ".f64"), "w") do io
        for p in particles; write(io, Float64.(p)); end
    end
end
first_plane(a) = ndims(a) == 2 ? a[:, 1] : a[:, :, 1]

function begin_step(simvars)
    empty!(scal); push!(scal, Float64(simvars.dt), Float64(simvars.t))
end
function after_update(mesh)
    put("fleck", mesh.fleck); put("beta", mesh.beta); put("bee", mesh.bee)
    put("sigma_a", first_plane(mesh.sigma_a)); put("sigma_s", first_plane(mesh.sigma_s))
end
function source_begin(particles)
    mode[] = :source; base[] = length(particles); empty!(uni); empty!(ex)
end
function source_end(mesh, particles)
    dump_ragged("source.uni", uni, length(particles) - base[])
    dump_particles("after_source", particles)
    put("emittedenergy", mesh.emittedenergy)
    push!(scal, Float64(mesh.totalenergy))
end
function transport_begin(particles)
    mode[] = :transport; empty!(uni); empty!(ex)
end
function transport_end(mesh, particles)
    dump_ragged("transport.uni", uni, length(particles)); dump_ragged("transport.exp", ex, length(particles))
    dump_particles("after_transport", particles)
    put("energydep", mesh.energydep)
    push!(scal, Float64(mesh.lostenergy))
end
function tally_end(mesh, particles)
    dump_particles("after_clean", particles)
    put("temp", mesh.temp); put("matenergydens", mesh.matenergydens); put("radenergydens", mesh.radenergydens)
    push!(scal, Float64(mesh.totalenergydep))
    put("scalars", scal)                     # [dt, t, totalenergy, lostenergy after transport, totalenergydep]
    step[] += 1
end
end # module REC

function patch(src, pairs...)
    for (a, b) in pairs
        occursin(a, src) || error("record_reference.jl: the reference no longer contains `$a` — update the patch list")
        src = replace(src, a => b)
    end
    src
end

function main_record(args)
    length(args) >= 3 || error("usage: julia record_reference.jl <reference checkout> <deck.txt> <out dir> [steps]")
    ref, deck, out = abspath(args[1]), abspath(args[2]), abspath(args[3])
    REC.nsteps[] = length(args) >= 4 ? parse(Int, args[4]) : 3
    mkpath(out); REC.outdir[] = out
    tmp = mktempdir()
    cp(joinpath(ref, "src"), joinpath(tmp, "src"))
    rd(f) = read(joinpath(tmp, "src", f), String)
    wr(f, s) = (chmod(joinpath(tmp, "src", f), 0o644); write(joinpath(tmp, "src", f), s))
    wr("imc_sourcing.jl", patch(rd("imc_sourcing.jl"), "rand(precision)" => "Main.REC.srand(particles, precision)"))
    wr("imc_transport.jl", patch(rd("imc_transport.jl"),
        "randexp(precision)" => "Main.REC.trandexp(precision)", "randexp()" => "Main.REC.trandexp()",
        "rand(precision)" => "Main.REC.trand(precision)",
        "for particle in eachindex(particles)" => "for particle in eachindex(particles); Main.REC.cur[] = particle"))
    wr("MixedPrecisionIMC.jl", patch(rd("MixedPrecisionIMC.jl"),
        "while simvars.t <= simvars.t_end" => "while simvars.t <= simvars.t_end && Main.REC.more()",
        "Update.update(inputs, mesh, simvars)" => "Main.REC.begin_step(simvars); Update.update(inputs, mesh, simvars); Main.REC.after_update(mesh)",
        "Sourcing.sourcing(mesh, simvars, particles)" => "Main.REC.source_begin(particles); Sourcing.sourcing(mesh, simvars, particles); Main.REC.source_end(mesh, particles)",
        "Transport.MC_RW(mesh, simvars, rwvars, particles)" => "(Main.REC.transport_begin(particles); Transport.MC_RW(mesh, simvars, rwvars, particles); Main.REC.transport_end(mesh, particles))",
        "Transport.MC(mesh, simvars, particles)" => "(Main.REC.transport_begin(particles); Transport.MC(mesh, simvars, particles); Main.REC.transport_end(mesh, particles))",
        "Transport.MC2D(mesh, simvars, particles)" => "Main.REC.transport_begin(particles); Transport.MC2D(mesh, simvars, particles); Main.REC.transport_end(mesh, particles)",
        "Tally.tally(inputs, mesh, simvars, particles)" => "Tally.tally(inputs, mesh, simvars, particles); Main.REC.tally_end(mesh, particles)",
        "Output.plotting(inputs, mesh, simvars)" => "nothing"))
    cp(deck, joinpath(out, "deck.txt"); force=true)
    empty!(ARGS); push!(ARGS, deck)                                 # loading the module runs main(ARGS) (MixedPrecisionIMC.jl:224-226)
    include(joinpath(tmp, "src", "MixedPrecisionIMC.jl"))
    inputs = Base.invokelatest(getfield(Main, :MixedPrecisionIMC).Input.readInputs, deck)
    open(joinpath(out, "manifest.json"), "w") do io
        print(io, "{\"producer\": \"MixedPrecisionIMC.jl under Julia ", VERSION, "\", \"precision\": \"", uppercase(string(inputs["PRECISION"])),
              "\", \"geometry\": \"", inputs["GEOMETRY"], "\", \"steps\": ", REC.step[], "}\n")
    end
    println("recorded ", REC.step[], " steps into ", out)
end

main_record(ARGS)
