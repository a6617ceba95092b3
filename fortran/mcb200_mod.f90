! mcb200_mod.f90 -- ISO_C_BINDING interface to libmocassin_b200.so (include/mcb200.h).
!
! This file is the binding a maintainer of the reference (rwesson/mocassin) adds to
! source/ so that iterateMC (source/iteration_mod.f90) drives the B200 transport instead
! of `call energyPacketDriver`.  It is shipped as source: the build image has no Fortran
! compiler, so it is compiled where the reference is (add it to SOURCES in the Makefile,
! link with -L<repo>/mocassin_b200 -lmocassin_b200).  See INTEGRATION.md for the edits to
! iteration_mod.f90.
!
! Every interface below is a 1:1 declaration of a C entry point; arrays are passed with
! c_loc of the reference's own allocatable components (their layout is already what the
! library expects: column major, cell index fastest, 1-based indices inside `active`).
module mcb200_mod
    use iso_c_binding
    implicit none

    type, bind(C) :: mcb200_config
        integer(c_int32_t) :: nGrids, nbins, nStars, nAngleBins, totAngleBinsTheta, totAngleBinsPhi, nLines
        integer(c_int32_t) :: lgDust, lgGas, lgSymmetricXYZ, lgIsotropic, lgPlaneIonization, lgDebug, &
             & lgMultistars, lgMultiDustChemistry
        integer(c_int32_t) :: nSpeciesMax, nSizes, nDustComp
        real(c_float)      :: dTheta, dPhi, R_out, ionEdge1
    end type mcb200_config

    type, bind(C) :: mcb200_counters
        integer(c_int64_t) :: nPackets, nAbs, nSca, trapped, nLinePackets, nDropped, nSegments, &
             & nFlights, nEscaped, nEarlyEscaped
        real(c_double)     :: Qphot, kernel_ms, total_ms
        integer(c_int64_t) :: nLaunches, nWaves
        real(c_double)     :: fly_ms
    end type mcb200_counters

    type(c_ptr), save :: mcb_ctx = c_null_ptr

    interface
       integer(c_int) function mcb200_create(ctx, device, rank, nranks, seed) bind(C, name="mcb200_create")
         import; type(c_ptr), intent(out) :: ctx
         integer(c_int32_t), value :: device, rank, nranks; integer(c_int64_t), value :: seed
       end function
       integer(c_int) function mcb200_destroy(ctx) bind(C, name="mcb200_destroy")
         import; type(c_ptr), value :: ctx
       end function
       type(c_ptr) function mcb200_last_error(ctx) bind(C, name="mcb200_last_error")
         import; type(c_ptr), value :: ctx
       end function
       integer(c_int) function mcb200_set_config(ctx, cfg) bind(C, name="mcb200_set_config")
         import; type(c_ptr), value :: ctx; type(mcb200_config), intent(in) :: cfg
       end function
       integer(c_int) function mcb200_set_grid(ctx, iG, nx, ny, nz, nCells, motherP, xAxis, yAxis, zAxis, active) &
            & bind(C, name="mcb200_set_grid")
         import; type(c_ptr), value :: ctx, xAxis, yAxis, zAxis, active
         integer(c_int32_t), value :: iG, nx, ny, nz, nCells, motherP
       end function
       integer(c_int) function mcb200_set_spectra(ctx, nuArray, gSca, inSpectrumProbDen) bind(C, name="mcb200_set_spectra")
         import; type(c_ptr), value :: ctx, nuArray, gSca, inSpectrumProbDen
       end function
       integer(c_int) function mcb200_set_stars(ctx, starPosition, starIndeces) bind(C, name="mcb200_set_stars")
         import; type(c_ptr), value :: ctx, starPosition, starIndeces
       end function
       integer(c_int) function mcb200_set_viewpoints(ctx, pTheta, pPhi, vTheta, vPhi) bind(C, name="mcb200_set_viewpoints")
         import; type(c_ptr), value :: ctx, pTheta, pPhi, vTheta, vPhi
       end function
       integer(c_int) function mcb200_set_dust_species(ctx, nSpeciesPart, grainAbun, dustComPoint, TdustSublime, nSpecies) &
            & bind(C, name="mcb200_set_dust_species")
         import; type(c_ptr), value :: ctx, nSpeciesPart, grainAbun, dustComPoint, TdustSublime
         integer(c_int32_t), value :: nSpecies
       end function
       integer(c_int) function mcb200_set_opacity(ctx, iG, opacity, scaOpac) bind(C, name="mcb200_set_opacity")
         import; type(c_ptr), value :: ctx, opacity, scaOpac; integer(c_int32_t), value :: iG
       end function
       integer(c_int) function mcb200_set_pdfs(ctx, iG, recPDF, dustPDF, totalLines, linePDF) bind(C, name="mcb200_set_pdfs")
         import; type(c_ptr), value :: ctx, recPDF, dustPDF, totalLines, linePDF; integer(c_int32_t), value :: iG
       end function
       integer(c_int) function mcb200_set_dust_state(ctx, iG, Tdust, dustAbunIndex) bind(C, name="mcb200_set_dust_state")
         import; type(c_ptr), value :: ctx, Tdust, dustAbunIndex; integer(c_int32_t), value :: iG
       end function
       integer(c_int) function mcb200_zero_estimators(ctx) bind(C, name="mcb200_zero_estimators")
         import; type(c_ptr), value :: ctx
       end function
       integer(c_int) function mcb200_transport(ctx, iStar, nPacketsGlobal, deltaE, counters) bind(C, name="mcb200_transport")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iStar
         integer(c_int64_t), value :: nPacketsGlobal; real(c_float), value :: deltaE
         type(mcb200_counters), intent(out) :: counters
       end function
       integer(c_int) function mcb200_transport_diffuse(ctx, gpLoc, cellLoc, nPacketsGlobal, deltaE, counters) &
            & bind(C, name="mcb200_transport_diffuse")
         import; type(c_ptr), value :: ctx, cellLoc; integer(c_int32_t), value :: gpLoc
         integer(c_int64_t), value :: nPacketsGlobal; real(c_float), value :: deltaE
         type(mcb200_counters), intent(out) :: counters
       end function
       integer(c_int) function mcb200_set_res_line_packets(ctx, iG, resLinePackets) bind(C, name="mcb200_set_res_line_packets")
         import; type(c_ptr), value :: ctx, resLinePackets; integer(c_int32_t), value :: iG
       end function
       integer(c_int) function mcb200_transport_reslines(ctx, iStar, deltaE, counters) bind(C, name="mcb200_transport_reslines")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iStar; real(c_float), value :: deltaE
         type(mcb200_counters), intent(out) :: counters
       end function
       integer(c_int) function mcb200_fetch_plane_distribution(ctx, planeIonDistribution) &
            & bind(C, name="mcb200_fetch_plane_distribution")
         import; type(c_ptr), value :: ctx, planeIonDistribution
       end function
       integer(c_int) function mcb200_tally_buffer(ctx, iG, which, devPtr, count) bind(C, name="mcb200_tally_buffer")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iG, which
         type(c_ptr), intent(out) :: devPtr; integer(c_int64_t), intent(out) :: count
       end function
       ! writeContCube's frequency sum per cell and viewing angle (output_mod.f90:2762-2772):
       ! contI(0:nCells, 0:nAngleBins), raw sums
       integer(c_int) function mcb200_fetch_contcube(ctx, iG, contI) bind(C, name="mcb200_fetch_contcube")
         import; type(c_ptr), value :: ctx, contI; integer(c_int32_t), value :: iG
       end function
       ! opacity(cells(r), 1:nbins) for a short list of cells: the rows writeTauNu reads along its rays
       integer(c_int) function mcb200_get_opacity_rows(ctx, iG, nWanted, cells, rows) bind(C, name="mcb200_get_opacity_rows")
         import; type(c_ptr), value :: ctx, cells, rows; integer(c_int32_t), value :: iG, nWanted
       end function
       integer(c_int) function mcb200_reduce(ctx) bind(C, name="mcb200_reduce")
         import; type(c_ptr), value :: ctx
       end function
       ! the library's own NCCL communicator: rank 0 makes the 128-byte id, the host broadcasts it
       ! (MPI_BCAST(id, 128, MPI_BYTE, 0, ...)), every rank joins; mcb200_exchange then replaces the
       ! MPI_ALLREDUCE block iteration_mod.f90:564,627,649,653,659
       integer(c_int) function mcb200_comm_unique_id(ctx, id128) bind(C, name="mcb200_comm_unique_id")
         import; type(c_ptr), value :: ctx, id128
       end function
       integer(c_int) function mcb200_comm_init(ctx, id128) bind(C, name="mcb200_comm_init")
         import; type(c_ptr), value :: ctx, id128
       end function
       integer(c_int) function mcb200_comm_destroy(ctx) bind(C, name="mcb200_comm_destroy")
         import; type(c_ptr), value :: ctx
       end function
       integer(c_int) function mcb200_exchange(ctx) bind(C, name="mcb200_exchange")
         import; type(c_ptr), value :: ctx
       end function
       integer(c_int) function mcb200_exchange_info(ctx, bytesSent, sparseGrids, ncclVersion) &
            & bind(C, name="mcb200_exchange_info")
         import; type(c_ptr), value :: ctx
         integer(c_int64_t), intent(out) :: bytesSent; integer(c_int32_t), intent(out) :: sparseGrids, ncclVersion
       end function
       integer(c_int) function mcb200_fetch_estimators(ctx, iG, Jste, escapedPackets, Jdif, linePackets) &
            & bind(C, name="mcb200_fetch_estimators")
         import; type(c_ptr), value :: ctx, Jste, escapedPackets, Jdif, linePackets; integer(c_int32_t), value :: iG
       end function
       ! K1: opacity, scaOpac, absOpac of grid iG on the device from the band list of addOpacity's
       ! inOpacity calls (ionization_mod.f90:349-484) and the dust loop (iteration_mod.f90:166-227);
       ! den(0:nCells, nSpeciesDen) = ionDen*elemAbun*Hden per species column, ff1 = FFOpacity(1) per cell
       integer(c_int) function mcb200_assemble_opacity(ctx, iG, nBands, bandSpecies, bandOff, bandLow, bandHigh, &
            & nSpeciesDen, den, ff1, Ndust, Tdust, dustAbunIndex, grainWeight, dustScaXsecP, dustAbsXsecP, nSpeciesTot) &
            & bind(C, name="mcb200_assemble_opacity")
         import; type(c_ptr), value :: ctx, bandSpecies, bandOff, bandLow, bandHigh, den, ff1
         type(c_ptr), value :: Ndust, Tdust, dustAbunIndex, grainWeight, dustScaXsecP, dustAbsXsecP
         integer(c_int32_t), value :: iG, nBands, nSpeciesDen, nSpeciesTot
       end function
       integer(c_int) function mcb200_get_opacity(ctx, iG, opacity, scaOpac, absOpac) bind(C, name="mcb200_get_opacity")
         import; type(c_ptr), value :: ctx, opacity, scaOpac, absOpac; integer(c_int32_t), value :: iG
       end function
       ! K8: nPhotoSte/heatSte (and Dif) per (cell, band) for updateCell / thermBalance (update_mod.f90:170-262, 1160-1214)
       integer(c_int) function mcb200_photo_integrals(ctx, iG, nBands, bandOff, bandLow, bandHigh, nPhotoSte, heatSte, &
            & nPhotoDif, heatDif) bind(C, name="mcb200_photo_integrals")
         import; type(c_ptr), value :: ctx, bandOff, bandLow, bandHigh, nPhotoSte, heatSte, nPhotoDif, heatDif
         integer(c_int32_t), value :: iG, nBands
       end function
       integer(c_int) function mcb200_fetch_qphot_counts(ctx, counts) bind(C, name="mcb200_fetch_qphot_counts")
         import; type(c_ptr), value :: ctx, counts
       end function
       integer(c_int) function mcb200_len_unit(ctx, iG, lenUnit) bind(C, name="mcb200_len_unit")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iG; real(c_double), intent(out) :: lenUnit
       end function
       ! Jste(iCell,:) of the cells one rank owns (iCell = firstCell, firstCell+cellStride, ...), compact
       integer(c_int) function mcb200_fetch_estimators_cells(ctx, iG, firstCell, cellStride, Jste, Jdif, nCellsOut) &
            & bind(C, name="mcb200_fetch_estimators_cells")
         import; type(c_ptr), value :: ctx, Jste, Jdif; integer(c_int32_t), value :: iG, firstCell, cellStride
         integer(c_int64_t), intent(out) :: nCellsOut
       end function
       ! how the last mcb200_exchange merged the J tallies (1 all-reduce, 2 reduce-scatter/all-gather, 3 peer-memory kernel)
       integer(c_int) function mcb200_exchange_path(ctx, path, why, whyLen, phaseMs) bind(C, name="mcb200_exchange_path")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), intent(out) :: path
         character(kind=c_char), dimension(*), intent(out) :: why; integer(c_int64_t), value :: whyLen
         type(c_ptr), value :: phaseMs       ! c_loc of 4 doubles, or c_null_ptr
       end function
       ! 64-bit position-sensitive checksum of a device-resident estimator (0 Jste, 1 escapedPackets, 2 Jdif, 3 linePackets)
       integer(c_int) function mcb200_checksum(ctx, iG, which, checksum) bind(C, name="mcb200_checksum")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iG, which; integer(c_int64_t), intent(out) :: checksum
       end function
       ! name is a C string: pass e.g. "wavefront"//c_null_char
       integer(c_int) function mcb200_set_option(ctx, name, value) bind(C, name="mcb200_set_option")
         import; type(c_ptr), value :: ctx; character(kind=c_char), dimension(*), intent(in) :: name
         integer(c_int64_t), value :: value
       end function
       ! escapedPackets sparsely (only its non-zero entries cross PCIe; the array must be zero elsewhere,
       ! as iterateMC leaves it at iteration_mod.f90:466-470), alone or together with the dense Jste
       integer(c_int) function mcb200_fetch_escaped_sparse(ctx, iG, escapedPackets, clearPrevious, nNonZero) &
            & bind(C, name="mcb200_fetch_escaped_sparse")
         import; type(c_ptr), value :: ctx, escapedPackets; integer(c_int32_t), value :: iG, clearPrevious
         integer(c_int64_t), intent(out) :: nNonZero
       end function
       integer(c_int) function mcb200_fetch_estimators_sparse(ctx, iG, Jste, escapedPackets, clearPrevious, nNonZero) &
            & bind(C, name="mcb200_fetch_estimators_sparse")
         import; type(c_ptr), value :: ctx, Jste, escapedPackets; integer(c_int32_t), value :: iG, clearPrevious
         integer(c_int64_t), intent(out) :: nNonZero
       end function
       integer(c_int) function mcb200_set_xsec(ctx, xSecArray, nXsec) bind(C, name="mcb200_set_xsec")
         import; type(c_ptr), value :: ctx, xSecArray; integer(c_int64_t), value :: nXsec
       end function
       ! dust-only closure: getDustT / updateCell (update_mod.f90:308-334,1836-1945), setDustPDF (emission_mod.f90:1313-1387)
       integer(c_int) function mcb200_set_dust_tables(ctx, widFlx, grainWeight, dustAbsXsecP, nSpecies, dustEmIntegral, nTemps) &
            & bind(C, name="mcb200_set_dust_tables")
         import; type(c_ptr), value :: ctx, widFlx, grainWeight, dustAbsXsecP, dustEmIntegral
         integer(c_int32_t), value :: nSpecies, nTemps
       end function
       integer(c_int) function mcb200_dust_update(ctx, iG, XHILimit, Tdust, lgConverged, nConverged) &
            & bind(C, name="mcb200_dust_update")
         import; type(c_ptr), value :: ctx, Tdust, lgConverged; integer(c_int32_t), value :: iG
         real(c_float), value :: XHILimit; integer(c_int64_t), intent(out) :: nConverged
       end function
       integer(c_int) function mcb200_dust_pdf(ctx, iG, dustPDF) bind(C, name="mcb200_dust_pdf")
         import; type(c_ptr), value :: ctx, dustPDF; integer(c_int32_t), value :: iG
       end function
       integer(c_int) function mcb200_reduce_range(ctx, iG, nu0, nu1) bind(C, name="mcb200_reduce_range")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iG, nu0, nu1
       end function
       ! sparse exchange of the escape counts: compact -> all-gather -> scatter (see include/mcb200.h)
       integer(c_int) function mcb200_escaped_compact(ctx, iG, set, devList, nEntries) bind(C, name="mcb200_escaped_compact")
         import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: iG, set
         type(c_ptr), intent(out) :: devList; integer(c_int64_t), intent(out) :: nEntries
       end function
       integer(c_int) function mcb200_escaped_scatter(ctx, iG, set, devList, nEntries) bind(C, name="mcb200_escaped_scatter")
         import; type(c_ptr), value :: ctx, devList; integer(c_int32_t), value :: iG, set
         integer(c_int64_t), value :: nEntries
       end function
       ! head of writeSED (output_mod.f90:2561-2568): SED(1:nbins,0:nAngleBins) raw sums over cells and grids
       integer(c_int) function mcb200_fetch_sed(ctx, SED, counts) bind(C, name="mcb200_fetch_sed")
         import; type(c_ptr), value :: ctx, SED, counts
       end function
    end interface

contains

    ! the reference's error behaviour: print and stop (e.g. photon_mod.f90:826-832)
    subroutine mcb_check(rc, where)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: where
        character(kind=c_char), pointer :: msg(:)
        integer :: i
        if (rc == 0) return
        call c_f_pointer(mcb200_last_error(mcb_ctx), msg, [512])
        write(*, '(a)', advance='no') '! '//where//': '
        do i = 1, 512
           if (msg(i) == c_null_char) exit
           write(*, '(a)', advance='no') msg(i)
        end do
        print*, ' [mcb200 status ', rc, ']'
        stop
    end subroutine mcb_check

end module mcb200_mod
