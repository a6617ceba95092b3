"""Recipe: translate the reference's hot-path Fortran, where it lies under
/root/reference/source, into oracle/_ref/mocassin_ref.py and mocassin_ref_aux.py
(TEST INFRASTRUCTURE).

    python -m oracle.f90ref.build_ref [--force]

Outputs only into oracle/_ref/ (git-ignored).  Nothing of the reference is copied into the
repository; the generated file is a build artefact like a compiled object.

What is translated (executable code):
    photon_mod.f90        whole module (energyPacketDriver and its internal procedures
                          energyPacketRun, initPhotonPacket, getNu, getNu2, newPhotonPacket,
                          pathSegment, hg)
    vector_mod.f90        whole module (type vector, its operators, randomUnitVector)
    interpolation_mod.f90 locate
    constants_mod.f90     parameters
What is only read for its declarations (module variables and derived types):
    common_mod.f90, continuum_mod.f90, grid_mod.f90, pathIntegration_mod.f90

Second target, mocassin_ref_aux.py -- the callers either side of the transport
(SURVEY.md 8a K1 and 8f.1), translated leniently: statements outside the pinned procedures
that the translator does not cover become calls that raise if they are ever reached.
    ionization_mod.f90    ionizationDriver, eDenSum, addOpacity (+ putOpacity, inOpacity)   strict
    continuum_mod.f90     getFlux                                         strict
    emission_mod.f90      emissionDriver with its internal setDustPDF     setDustPDF strict
    update_mod.f90        updateCell with its internal getDustT           getDustT strict
    output_mod.f90        writeSED (list-directed WRITEs are recorded, OPEN/CLOSE do nothing)  strict
    grid_mod.f90          writeGrid (grid0-3.out, dustGrid.out, photoSource.out records)        strict
    hydro_mod.f90         getOuterShell                                                     strict
    update_mod.f90        lines 168-269 of updateCell (photo-ionisation rates nPhotoSte/nPhotoDif) and
                          lines 1123-1234 of thermBalance (photo-ionisation heating), each as a
                          synthetic subroutine with the host's declarations                 strict
    iteration_mod.f90     lines 106-230 of iterateMC (the opacity block: ionizationDriver over the
                          cells + the dust contribution to scaOpac/absOpac/opacity) as a
                          synthetic subroutine with iterateMC's declarations                strict
(the one untranslated statement inside each of setDustPDF / getDustT is the call of the
quantum-heating / resonance-line-heating routine, in a branch the dust-only path never takes)
"""
from __future__ import annotations

import os
import sys

from . import f90py

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), '_ref')
OUT = os.path.join(OUT_DIR, 'mocassin_ref.py')
OUT_AUX = os.path.join(OUT_DIR, 'mocassin_ref_aux.py')
REF_ROOT = os.environ.get('MOCASSIN_REFERENCE', '/root/reference')

# (file, declarations only?, procedures to keep or None for all)
SOURCES = [
    ('constants_mod.f90', False, None),
    ('vector_mod.f90', False, None),
    ('interpolation_mod.f90', False, {'locate'}),
    ('common_mod.f90', True, None),
    ('continuum_mod.f90', True, None),
    ('grid_mod.f90', True, None),
    ('pathIntegration_mod.f90', True, None),
    ('photon_mod.f90', False, None),
]
# instrumentation planted by the translator (calls into rt.loop_hook / rt.proc_hook):
LOOP_HOOKS = {'energypacketdriver.iphot',    # one trip per energy packet (photon_mod.f90:92, :201)
              'pathsegment.j'}               # one trip per cell crossing (photon_mod.f90:1194)
PROC_HOOKS = {'energypacketrun'}             # one call per packet generation


# (file, declarations only?, procedures to keep, internal procedures to keep)
AUX_SOURCES = [
    ('constants_mod.f90', False, None, None),
    ('vector_mod.f90', False, None, None),       # the vector operators (integratePathTauNu: rVec + dlSmall*vHat)
    ('interpolation_mod.f90', False, {'locate', 'linearmap', 'sortup'}, None),
    ('common_mod.f90', True, None, None),
    # module xSec_mod; BHmie (COMPLEX arithmetic, statement functions); the gas cross-section stack
    ('ph_mod.f90', False, {'bhmie', 'getqs', 'initxsecarray', 'phfitel', 'phfithion', 'powlawxsec', 'makeopacity'}, None),
    # module elements_mod: level energies, shell / continuum pointers
    ('hydro_mod.f90', False, {'getoutershell', 'makehydro', 'setshells', 'limitshell', 'setpointers'}, None),
    ('grid_mod.f90', False, {'writegrid', 'setstarposition', 'getvolume'}, None),
    ('composition_mod.f90', True, None, None),
    ('set_input_mod.f90', True, None, None),
    ('continuum_mod.f90', False, {'getflux', 'setprobden'}, None),
    ('ionization_mod.f90', False, {'ionizationdriver', 'edensum', 'addopacity'}, None),
    ('emission_mod.f90', False, {'emissiondriver'}, {'setdustpdf'}),
    ('update_mod.f90', False, {'updatecell'}, {'getdustt'}),
    ('pathIntegration_mod.f90', False, {'integratepathtaunu'}, None),
    ('output_mod.f90', False, {'writesed', 'writecontcube', 'writetaunu'}, None),
]
# Statement ranges of procedures that cannot be run as a whole (iterateMC is the entire Lucy
# iteration, MPI included; updateCell's gas branch is the whole ionisation/thermal solver),
# wrapped -- with the host procedure's own declarations -- into synthetic subroutines.
# dict(file, name, args, decls = line ranges of declarations, body = line ranges of statements,
#      glue_decls / glue_start / glue_end = the only lines that are ours: the dummies of the wrapper,
#      the copy of the dummies into the host's locals, and the copy of a local result into the
#      dummies, guards = {line: expected start} against a changed file)
AUX_SLICES = [
    # the opacity block of iterateMC: ionizationDriver over all cells, the all-reduce, and the dust
    # contribution to scaOpac/absOpac/opacity
    dict(file='iteration_mod.f90', name='opacity_block', args='grid',
         decls=[(20, 20), (24, 24), (37, 77)], body=[(106, 230)], glue_decls=[], glue_end=[],
         guards={106: 'icell = 0', 108: 'do ig = 1, ngrids', 230: 'end do'}),
    # updateCell: number of stellar/diffuse photo-ionisations per (element, ion)
    dict(file='update_mod.f90', name='photo_rates', args='grid, xp, yp, zp, outste, outdif',
         decls=[(18, 86)], body=[(90, 98), (168, 269)],
         glue_decls=['real, intent(out) :: outste(nelements, nstages), outdif(nelements, nstages)'],
         glue_end=['outste = nphotoste', 'outdif = nphotodif'],
         guards={92: 'cellp = grid%active', 168: 'nphotoste = 1.e-20', 269: 'end do'}),
    # thermBalance: heating by photo-ionisation, heatSte/heatDif
    dict(file='update_mod.f90', name='photo_heat', args='grid, xp, yp, zp, outste, outdif',
         decls=[(18, 86), (936, 959)], body=[(90, 98), (1123, 1234)],
         glue_decls=['real, intent(out) :: outste, outdif'],
         glue_end=['outste = heatste', 'outdif = heatdif'],
         guards={936: 'real, intent(out) :: heatint', 1123: 'heatste = 0.', 1234: 'end do'}),
    # fillGrid: automatic axes (symmetric octant or full cube), the geometric corrections, and the
    # masking of the cells that lie inside another grid
    dict(file='grid_mod.f90', name='fill_axes', args='grid', decls=[(493, 498)], body=[(530, 601)],
         glue_decls=[], glue_end=[], guards={493: 'type(grid_type), dimension(:),intent(inout) :: grid', 530: 'if (.not.lgdfile) then',
                                             538: 'grid(1)%xaxis(i) = grid(1)%xaxis(i) * rnx', 601: 'end if'}),
    dict(file='grid_mod.f90', name='fill_mask', args='grid', decls=[(493, 498)], body=[(807, 816), (833, 833), (835, 889)],
         glue_decls=[], glue_end=[], guards={807: 'do ig = 1, ngrids', 809: 'grid(ig)%geocorrx =', 816: 'end if', 833: 'end do', 835: 'if (ngrids>1) then',
                                             862: 'grid(ig)%active(i,j,k) = -jg', 889: 'end if'}),
    # setMotherGrid: which cells are active and how they are numbered (radius test against R_in /
    # R_out, then a running count over cells that hold gas or dust)
    dict(file='grid_mod.f90', name='active_cells', args='grid, in_hden, in_ndust, in_ytop',
         decls=[(901, 901), (904, 938)], body=[(1226, 1294)],
         glue_decls=['real, intent(in) :: in_hden(grid%nx, grid%ny, grid%nz), in_ndust(grid%nx, grid%ny, grid%nz)',
                     'integer, intent(in) :: in_ytop'],
         glue_start=['allocate(hdentemp(1:grid%nx, 1:grid%ny, 1:grid%nz))', 'allocate(ndusttemp(1:grid%nx, 1:grid%ny, 1:grid%nz))',
                     'hdentemp = in_hden', 'ndusttemp = in_ndust', 'ytop = in_ytop', 'grid%active = 1'],
         glue_end=[],
         guards={901: 'type(grid_type), intent(inout) :: grid', 938: 'character(len=40)', 1226: 'grid%ncells = 0',
                 1235: 'radius = 1.e10*sqrt(', 1294: 'end do'}),
    # initCartesianGrid: the ionisation thresholds inside the frequency range, the gas-only frequency
    # mesh (series edges, thresholds, logarithmic fill, sortUp) and widFlx
    dict(file='grid_mod.f90', name='gas_nu_mesh', args='', decls=[(31, 47)], body=[(132, 175), (215, 258), (333, 338)],
         glue_decls=[], glue_start=['allocate(nuarray(1:nbins))', 'allocate(widflx(1:nbins))'], glue_end=[],
         guards={31: 'integer :: err, ios', 132: 'seriesedge = (/0.0069', 175: 'call sortup(ionedge(1:nedges))',
                 216: 'if (numin<radio4p9ghz) then', 258: 'call sortup(nuarray)', 333: 'widflx(1) = nuarray(2)-nuarray(1)',
                 338: 'widflx(nbins) = nuarray(nbins)-nuarray(nbins-1)'}),
    # setMotherGrid: the ionisation state every active cell starts from
    dict(file='grid_mod.f90', name='initial_ions', args='grid, in_ytop', decls=[(901, 901), (904, 938)], body=[(1564, 1607)],
         glue_decls=['integer, intent(in) :: in_ytop'], glue_start=['ytop = in_ytop'], glue_end=[],
         guards={1564: 'h0in = 1.e-5', 1607: 'end do'}),
    # setSubGrids: the sub-grid list and one sub-grid's density file -- list-directed READs (units 71, 72 bound
    # by the harness), axes rescaled from normalised coordinates, the reference's own sanity stops, active cells
    dict(file='grid_mod.f90', name='sub_grid_read', args='grid, out_hden, out_ndust', decls=[(1807, 1836)],
         body=[(1847, 1857), (2028, 2232)],
         glue_decls=['real, intent(out) :: out_hden(:,:,:), out_ndust(:,:,:)'],
         glue_end=['out_hden = hdentemp', 'if (lgdust) out_ndust = ndusttemp', 'end do'],
         guards={1807: 'type(grid_type), dimension(:),intent(inout) :: grid', 1847: 'do ig = 2, ngrids',
                 1853: 'read(71, *) grid(ig)%motherp', 2029: 'open (unit= 72, file=dfileread', 2103: 'read(72, *) x,y,z, hdentemp(ix,iy,iz), ndusttemp(ix,iy,iz)',
                 2230: 'end do', 2232: 'close(72)'}),
    # initCartesianGrid: angular bins of the escape tallies and the viewing-angle pointer tables
    dict(file='grid_mod.f90', name='angle_tables', args='', decls=[], body=[(416, 468)],
         glue_decls=['integer :: i, err'], glue_end=[],
         guards={416: 'dtheta = pi/totanglebinstheta', 429: 'dphi = twopi/totanglebinsphi',
                 466: 'viewpointptheta(int(viewpointtheta(i)/dtheta)+1) = i', 468: 'end do'}),
    # makeDustXsec: trapezoid widths of the size grid and the normalisation of the grain weights
    dict(file='ph_mod.f90', name='grain_weights', args='', decls=[(810, 835)], body=[(986, 1012)],
         glue_decls=[], glue_start=['allocate(da(1:nsizes))', 'da = 0.'], glue_end=[],
         guards={986: 'if (nsizes>1) then', 1000: 'grainweight(ai) = (grainweight(ai)*da(ai))/normweight', 1012: 'end if'}),
    # dustEmissionInt (internal to dustDriver): the emission integrals getDustT inverts
    dict(file='dust_mod.f90', name='dust_emission_int', args='',
         decls=[(148, 152)], body=[(155, 181)], glue_decls=['integer :: err'], glue_end=[],
         guards={148: 'real :: bb', 155: 'allocate(dustemintegral(1:nspecies,1:nsizes,ntemps)', 170: 'bb = getflux(nuarray(i), real(nt), cshapeloc)',
                 181: 'dustemintegral = dustemintegral*hplanck*4.'}),
    # makeDustXsec: from efficiencies to cross-sections, the dust part of xSecArray, its pointer
    # tables and gSca, for one dust component (the tail of the icomp loop)
    dict(file='ph_mod.f90', name='dust_xsec_assembly', args='in_icomp, in_csca, in_cabs, in_gcos',
         decls=[(810, 835)], body=[(1456, 1538)],
         glue_decls=['integer, intent(in) :: in_icomp',
                     'real, intent(in) :: in_csca(nspecies, 0:nsizes, nbins), in_cabs(nspecies, 0:nsizes, nbins), '
                     'in_gcos(nspecies, 0:nsizes, nbins)'],
         glue_start=['icomp = in_icomp',
                     'allocate(csca(1:nspecies, 0:nsizes, 1:nbins))', 'allocate(cabs(1:nspecies, 0:nsizes, 1:nbins))',
                     'allocate(gcos(1:nspecies, 0:nsizes, 1:nbins))', 'allocate(ctsca(1:nbins))', 'allocate(ctabs(1:nbins))',
                     'allocate(norm(nbins))',
                     'csca = in_csca', 'cabs = in_cabs', 'gcos = in_gcos', 'ctsca = 0.', 'ctabs = 0.'],
         glue_end=[],
         guards={810: 'real :: normweight', 835: 'character(len=50) :: extinctionfile', 1456: 'do i = 1, nbins',
                 1460: 'csca(nspec,ai,i) = csca(nspec,ai,i)*pi', 1538: 'enddo'}),
]
# supplied by the harness: BoltGaunt (ionization_mod.f90:134-174) fills contBoltz/gauntFF from Gaunt
# factor tables; its only trace in the opacity is the free-free term of bin 1, which the
# oracle takes as an input (ff1), so the harness sets those arrays directly
AUX_EXTERNS = {'boltgaunt',
               # file readers and branches off the gas-deck path (initXSecArray, setPointers): supplied as no-ops
               'makecolliondata', 'makeaugerdata', 'readheireclines', 'setcompton', 'makedustxsec', 'phinit'}
# procedures that must translate completely, and the untranslated statements tolerated in them
AUX_STRICT = {'writegrid': 0, 'setstarposition': 0, 'getvolume': 0, 'writesed': 0, 'writecontcube': 0, 'writetaunu': 0, 'integratepathtaunu': 0, 'bhmie': 0, 'getqs': 0, 'dust_xsec_assembly': 0, 'dust_emission_int': 0, 'grain_weights': 0, 'angle_tables': 0, 'active_cells': 0, 'fill_axes': 0, 'fill_mask': 0, 'opacity_block': 0, 'photo_rates': 0, 'photo_heat': 0, 'getoutershell': 0, 'ionizationdriver': 0, 'edensum': 0, 'addopacity': 0, 'putopacity': 0, 'inopacity': 0, 'getflux': 0, 'setprobden': 0, 'locate': 0, 'linearmap': 0, 'sortup': 0, 'gas_nu_mesh': 0, 'initial_ions': 0, 'sub_grid_read': 0, 'initxsecarray': 0, 'phfitel': 0, 'phfithion': 0, 'powlawxsec': 0, 'makeopacity': 0, 'makehydro': 0, 'setshells': 0, 'limitshell': 0, 'setpointers': 0,
              'setdustpdf': 1,      # call qHeat (lgQHeat branch)
              'getdustt': 1}        # resLineHeating (gas + resonance-line transfer branch)


class ReferenceUnavailable(RuntimeError):
    pass


def source_dir():
    return os.path.join(REF_ROOT, 'source')


def available() -> bool:
    return os.path.exists(OUT) or os.path.isdir(source_dir())


def translate() -> str:
    mods = []
    for fn, spec_only, keep in SOURCES:
        with open(os.path.join(source_dir(), fn), errors='replace') as fh:
            unit = f90py.Unit(fh.read(), fn, spec_only=spec_only)
        for m in unit.modules:
            if keep is not None:
                m.procs = {k: v for k, v in m.procs.items() if k in keep}
            mods.append(m)
    gen = f90py.Gen(mods, loop_hooks=LOOP_HOOKS, proc_hooks=PROC_HOOKS)      # strict: every statement
    code = gen.generate('@@HEADER@@')
    header = ('# GENERATED by oracle/f90ref/build_ref.py from the Fortran sources under\n'
              f'# {source_dir()} -- a build artefact, do not edit, do not commit.\n'
              '# translator notes:\n' + ''.join(f'#   {w}\n' for w in gen.warnings))
    return code.replace('@@HEADER@@', header, 1)


def translate_aux() -> str:
    mods = []
    for fn, spec_only, keep, keep_internal in AUX_SOURCES:
        with open(os.path.join(source_dir(), fn), errors='replace') as fh:
            unit = f90py.Unit(fh.read(), fn, spec_only=spec_only, lenient=True)
        for m in unit.modules:
            if keep is not None:
                m.procs = {k: v for k, v in m.procs.items() if k in keep}
            if keep_internal is not None:
                for p in m.procs.values():
                    p.contains = {k: v for k, v in p.contains.items() if k in keep_internal}
            mods.append(m)
    for sl in AUX_SLICES:
        fn, name = sl['file'], sl['name']
        with open(os.path.join(source_dir(), fn), errors='replace') as fh:
            lines = fh.read().split('\n')
        for ln, want in sl['guards'].items():
            got = ' '.join(lines[ln - 1].split('!')[0].lower().split())
            if not got.startswith(want):
                raise f90py.TranslateError(f'{fn}:{ln}: expected {want!r}, found {got!r} (reference changed?)')
        text = ['module slice_' + name, 'contains', f'subroutine {name}({sl["args"]})']
        for a, b in sl['decls']:
            text += lines[a - 1:b]
        text += sl['glue_decls']
        text += sl.get('glue_start', [])
        for a, b in sl['body']:
            text += lines[a - 1:b]
        text += sl['glue_end'] + [f'end subroutine {name}', 'end module slice_' + name]
        unit = f90py.Unit('\n'.join(text), f'{fn}[{sl["body"]}]', lenient=True)
        mods.extend(unit.modules)
    gen = f90py.Gen(mods, lenient=True, externs=AUX_EXTERNS)
    code = gen.generate('@@HEADER@@')
    for name, allowed in AUX_STRICT.items():
        bad = gen.untranslated.get(name, [])
        if len(bad) > allowed:
            raise f90py.TranslateError(f'{name}: untranslated statements in a pinned procedure: {bad}')
    notes = [f'{k}: {x}' for k, v in gen.untranslated.items() for x in v]
    header = ('# GENERATED by oracle/f90ref/build_ref.py from the Fortran sources under\n'
              f'# {source_dir()} -- a build artefact, do not edit, do not commit.\n'
              '# statements that raise if reached (lenient translation):\n' + ''.join(f'#   {w}\n' for w in notes))
    return code.replace('@@HEADER@@', header, 1)


def _write(path, code):
    os.makedirs(OUT_DIR, exist_ok=True)
    compile(code, path, 'exec')
    tmp = path + '.tmp'
    with open(tmp, 'w') as fh:
        fh.write(code)
    os.replace(tmp, path)


def _stale_path(path, sources) -> bool:
    if not os.path.exists(path):
        return True
    if not os.path.isdir(source_dir()):
        return False            # the box without the reference uses the prebuilt file
    t = os.path.getmtime(path)
    deps = [os.path.join(source_dir(), f[0]) for f in sources]
    deps += [os.path.join(HERE, f) for f in ('f90py.py', 'build_ref.py')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, target: str = 'photon') -> str:
    path, sources, fn = (OUT, SOURCES, translate) if target == 'photon' else (OUT_AUX, AUX_SOURCES, translate_aux)
    if not force and not _stale_path(path, sources):
        return path
    if not os.path.isdir(source_dir()):
        raise ReferenceUnavailable(f'{source_dir()} not present and {path} not built')
    _write(path, fn())
    return path


if __name__ == '__main__':
    for tgt in ('photon', 'aux'):
        print(build(force='--force' in sys.argv, target=tgt))
