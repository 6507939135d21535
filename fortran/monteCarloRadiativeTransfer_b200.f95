! Replacement for Integrators/monteCarloRadiativeTransfer.f95: same module name, same public list
! (monteCarloRadiativeTransfer.f95:154-156), every procedure forwards to the C ABI of include/i3rc_b200.h.
! NOT COMPILED in the build environment (no Fortran compiler exists there); kept short so that it can be reviewed.
! Link the drivers against this module and libi3rc_b200.so instead of the reference integrator.
module monteCarloRadiativeTransfer
  use, intrinsic :: iso_c_binding
  use ErrorMessages,     only: ErrorMessage, stateIsFailure, setStateToFailure, setStateToWarning, &
                               setStateToSuccess, setStateToCompleteSuccess
  use RandomNumbers,     only: randomNumberSequence
  use scatteringPhaseFunctions, only: phaseFunctionTable, phaseFunction, getInfo_PhaseFunctionTable, getElement, &
                               getInfo_PhaseFunction, getPhaseFunctionCoefficients, getPhaseFunctionValues
  use opticalProperties, only: domain, getInfo_Domain, getOpticalPropertiesByComponent
  use surfaceProperties, only: surfaceDescription
  use monteCarloIllumination, only: photonStream, i3rc_photon_source, descriptorOf   ! fortran/monteCarloIllumination_b200.f95
  implicit none
  private

  type integrator
    private
    type(c_ptr) :: handle = c_null_ptr
    integer     :: nX = 0, nY = 0, nZ = 0, nComponents = 0, nDirections = 0
  end type integrator

  ! mirrors of the C structs (include/i3rc_b200.h)
  type, bind(c) :: i3rc_phase_table
    integer(c_int32_t) :: kind, n_entries
    type(c_ptr)        :: coef_offsets, coefs
    integer(c_int32_t) :: n_angles
    type(c_ptr)        :: angles, values
  end type
  type, bind(c) :: i3rc_params
    integer(c_int32_t) :: present
    real(c_float)      :: surfaceAlbedo
    integer(c_int32_t) :: minForwardTableSize, minInverseTableSize, numIntensityDirections
    type(c_ptr)        :: intensityMus, intensityPhis
    integer(c_int32_t) :: computeIntensity, useRayTracing, useRussianRoulette, useRussianRouletteForIntensity
    real(c_float)      :: zetaMin
    integer(c_int32_t) :: useHybridPhaseFunsForIntenCalcs
    real(c_float)      :: hybridPhaseFunWidth
    integer(c_int32_t) :: numOrdersOrigPhaseFunIntenCalcs, limitIntensityContributions
    real(c_float)      :: maxIntensityContribution
    integer(c_int32_t) :: surf_nx, surf_ny
    type(c_ptr)        :: surf_x, surf_y, surf_params
  end type
  ! (i3rc_photon_source is declared by the replacement module monteCarloIllumination, which fills it)

  interface
    integer(c_int) function i3rc_new_Integrator(nx, ny, nz, nc, xPos, yPos, zPos, totalExt, cumExt, ssa, pfIndex, h) bind(c)
      import; integer(c_int), value :: nx, ny, nz, nc
      real(c_float), intent(in) :: xPos(*), yPos(*), zPos(*), totalExt(*), cumExt(*), ssa(*)
      integer(c_int32_t), intent(in) :: pfIndex(*); type(c_ptr), intent(out) :: h
    end function
    integer(c_int) function i3rc_set_phase_table(h, comp, t) bind(c)
      import; type(c_ptr), value :: h; integer(c_int), value :: comp; type(i3rc_phase_table), intent(in) :: t
    end function
    integer(c_int) function i3rc_specifyParameters(h, p) bind(c)
      import; type(c_ptr), value :: h; type(i3rc_params), intent(in) :: p
    end function
    integer(c_int) function i3rc_computeRadiativeTransfer(h, src, seed, nseed) bind(c)
      import; type(c_ptr), value :: h; type(i3rc_photon_source), intent(in) :: src
      integer(c_int32_t), intent(in) :: seed(*); integer(c_int), value :: nseed
    end function
    integer(c_int) function i3rc_reportResults(h, mUp, mDown, mAbs, fUp, fDown, fAbs, prof, vol, mInt, inten) bind(c)
      import; type(c_ptr), value :: h, mUp, mDown, mAbs, fUp, fDown, fAbs, prof, vol, mInt, inten
    end function
    integer(c_int) function i3rc_copy_Integrator(src, h) bind(c)
      import; type(c_ptr), value :: src; type(c_ptr), intent(out) :: h
    end function
    subroutine i3rc_finalize_Integrator(h) bind(c)
      import; type(c_ptr), value :: h
    end subroutine
    integer(c_int) function i3rc_isReady_Integrator(h) bind(c)
      import; type(c_ptr), value :: h
    end function
    type(c_ptr) function i3rc_last_message(h) bind(c)
      import; type(c_ptr), value :: h
    end function
  end interface

  public :: integrator
  public :: new_Integrator, copy_Integrator, isReady_Integrator, finalize_Integrator, &
            specifyParameters, computeRadiativeTransfer, reportResults
contains
  ! ---- status mapping: 0 success, 1 warning, 2 failure (Code/ErrorMessages.f95:159-248)
  subroutine toStatus(rc, h, status, where)
    integer(c_int), intent(in) :: rc; type(c_ptr), intent(in) :: h
    type(ErrorMessage), intent(inout) :: status; character(len=*), intent(in) :: where
    character(len=256) :: text
    text = where // ": " // cString(i3rc_last_message(h))
    if (rc == 2) then; call setStateToFailure(status, trim(text))
    else if (rc == 1) then; call setStateToWarning(status, trim(text))
    else; call setStateToSuccess(status); end if
  end subroutine
  function cString(p) result(s)
    type(c_ptr), intent(in) :: p; character(len=256) :: s
    character(kind=c_char), pointer :: c(:); integer :: i
    s = ""; if (.not. c_associated(p)) return
    call c_f_pointer(p, c, (/ 256 /))
    do i = 1, 256; if (c(i) == c_null_char) exit; s(i:i) = c(i); end do
  end function

  ! ---- new_Integrator (monteCarloRadiativeTransfer.f95:162-254)
  function new_Integrator(atmosphere, status) result(new)
    type(domain), intent(in) :: atmosphere; type(ErrorMessage), intent(inout) :: status; type(integrator) :: new
    integer :: nX, nY, nZ, nC, i
    real, allocatable :: x(:), y(:), z(:), totalExt(:,:,:), cumExt(:,:,:,:), ssa(:,:,:,:)
    integer, allocatable :: pfIndex(:,:,:,:)
    type(phaseFunctionTable), allocatable :: tables(:)
    call getInfo_Domain(atmosphere, nX, nY, nZ, numberOfComponents = nC, status = status)
    if (stateIsFailure(status)) return
    allocate(x(nX+1), y(nY+1), z(nZ+1), totalExt(nX,nY,nZ), cumExt(nX,nY,nZ,nC), ssa(nX,nY,nZ,nC), pfIndex(nX,nY,nZ,nC), tables(nC))
    call getInfo_Domain(atmosphere, xPosition = x, yPosition = y, zPosition = z, status = status)
    call getOpticalPropertiesByComponent(atmosphere, totalExt, cumExt, ssa, pfIndex, tables, status)
    call toStatus(i3rc_new_Integrator(nX, nY, nZ, nC, x, y, z, totalExt, cumExt, ssa, pfIndex, new%handle), &
                  new%handle, status, "new_Integrator")
    if (stateIsFailure(status)) return
    new%nX = nX; new%nY = nY; new%nZ = nZ; new%nComponents = nC
    do i = 1, nC
      call passPhaseTable(new, i, tables(i), status)   ! forwardTables(i), monteCarloRadiativeTransfer.f95:92-93
    end do
  end function

  ! One phaseFunctionTable -> i3rc_phase_table: Legendre coefficients (getPhaseFunctionCoefficients) or the native
  ! angle/value pairs of a one-angle-set table (getInfo_PhaseFunction(nativeAngles), getPhaseFunctionValues).
  subroutine passPhaseTable(this, comp, table, status)
    type(integrator), intent(in) :: this; integer, intent(in) :: comp
    type(phaseFunctionTable), intent(in) :: table; type(ErrorMessage), intent(inout) :: status
    type(i3rc_phase_table) :: t
    type(phaseFunction) :: pf
    integer :: nEntries, e, nCoef, nAng, total
    integer(c_int32_t), allocatable, target :: offsets(:)
    real(c_float), allocatable, target :: coefs(:), angles(:), values(:,:)
    call getInfo_PhaseFunctionTable(table, nEntries = nEntries, status = status)
    allocate(offsets(nEntries+1)); offsets(1) = 0
    pf = getElement(1, table, status)
    call getInfo_PhaseFunction(pf, nCoef, nAng, status = status)
    if (nAng > 0) then                       ! tabulated, one angle set
      allocate(angles(nAng), values(nAng, nEntries))
      call getInfo_PhaseFunction(pf, nativeAngles = angles, status = status)
      do e = 1, nEntries
        pf = getElement(e, table, status); call getPhaseFunctionValues(pf, angles, values(:, e), status)
      end do
      t%kind = 2; t%n_entries = nEntries; t%n_angles = nAng; t%angles = c_loc(angles); t%values = c_loc(values)
    else                                     ! Legendre series, possibly of different lengths
      total = 0
      do e = 1, nEntries
        pf = getElement(e, table, status); call getInfo_PhaseFunction(pf, nCoef, nAng, status = status)
        total = total + nCoef; offsets(e+1) = total
      end do
      allocate(coefs(max(total, 1)))
      do e = 1, nEntries
        pf = getElement(e, table, status)
        call getPhaseFunctionCoefficients(pf, coefs(offsets(e)+1:offsets(e+1)), status)
      end do
      t%kind = 1; t%n_entries = nEntries; t%coef_offsets = c_loc(offsets); t%coefs = c_loc(coefs)
    end if
    call toStatus(i3rc_set_phase_table(this%handle, comp - 1, t), this%handle, status, "new_Integrator")
  end subroutine

  ! ---- specifyParameters (monteCarloRadiativeTransfer.f95:830-1069): optional arguments -> presence mask
  subroutine specifyParameters(thisIntegrator, surfaceAlbedo, surfaceBDRF, minForwardTableSize, minInverseTableSize, &
                               intensityMus, intensityPhis, computeIntensity, useRayTracing, useRussianRoulette,     &
                               useRussianRouletteForIntensity, zetaMin, useHybridPhaseFunsForIntenCalcs,             &
                               hybridPhaseFunWidth, numOrdersOrigPhaseFunIntenCalcs, limitIntensityContributions,    &
                               maxIntensityContribution, status)
    type(integrator), intent(inout) :: thisIntegrator
    real, optional, intent(in) :: surfaceAlbedo, zetaMin, hybridPhaseFunWidth, maxIntensityContribution
    type(surfaceDescription), optional, intent(in) :: surfaceBDRF
    integer, optional, intent(in) :: minForwardTableSize, minInverseTableSize, numOrdersOrigPhaseFunIntenCalcs
    real, dimension(:), optional, target, intent(in) :: intensityMus, intensityPhis
    logical, optional, intent(in) :: computeIntensity, useRayTracing, useRussianRoulette, useRussianRouletteForIntensity, &
                                     useHybridPhaseFunsForIntenCalcs, limitIntensityContributions
    type(ErrorMessage), intent(inout) :: status
    type(i3rc_params) :: p
    p%present = 0
    if (present(surfaceAlbedo))       then; p%present = ior(p%present, 2**0);  p%surfaceAlbedo = surfaceAlbedo; end if
    ! surfaceBDRF (bit 1): the albedo map of the surfaceDescription is passed through surf_x/surf_y/surf_params; it needs
    ! an accessor in module surfaceProperties (its components are private, surfaceProperties.f95:34-38).
    if (present(minForwardTableSize)) then; p%present = ior(p%present, 2**2);  p%minForwardTableSize = minForwardTableSize; end if
    if (present(minInverseTableSize)) then; p%present = ior(p%present, 2**3);  p%minInverseTableSize = minInverseTableSize; end if
    if (present(intensityMus))        then; p%present = ior(p%present, 2**4);  p%intensityMus = c_loc(intensityMus)
                                            p%numIntensityDirections = size(intensityMus); end if
    if (present(intensityPhis))       then; p%present = ior(p%present, 2**5);  p%intensityPhis = c_loc(intensityPhis); end if
    if (present(computeIntensity))    then; p%present = ior(p%present, 2**6);  p%computeIntensity = merge(1, 0, computeIntensity); end if
    if (present(useRayTracing))       then; p%present = ior(p%present, 2**7);  p%useRayTracing = merge(1, 0, useRayTracing); end if
    if (present(useRussianRoulette))  then; p%present = ior(p%present, 2**8);  p%useRussianRoulette = merge(1, 0, useRussianRoulette); end if
    if (present(useRussianRouletteForIntensity)) then
      p%present = ior(p%present, 2**9); p%useRussianRouletteForIntensity = merge(1, 0, useRussianRouletteForIntensity); end if
    if (present(zetaMin))             then; p%present = ior(p%present, 2**10); p%zetaMin = zetaMin; end if
    if (present(useHybridPhaseFunsForIntenCalcs)) then
      p%present = ior(p%present, 2**11); p%useHybridPhaseFunsForIntenCalcs = merge(1, 0, useHybridPhaseFunsForIntenCalcs); end if
    if (present(hybridPhaseFunWidth)) then; p%present = ior(p%present, 2**12); p%hybridPhaseFunWidth = hybridPhaseFunWidth; end if
    if (present(numOrdersOrigPhaseFunIntenCalcs)) then
      p%present = ior(p%present, 2**13); p%numOrdersOrigPhaseFunIntenCalcs = numOrdersOrigPhaseFunIntenCalcs; end if
    if (present(limitIntensityContributions)) then
      p%present = ior(p%present, 2**14); p%limitIntensityContributions = merge(1, 0, limitIntensityContributions); end if
    if (present(maxIntensityContribution)) then
      p%present = ior(p%present, 2**15); p%maxIntensityContribution = maxIntensityContribution; end if
    call toStatus(i3rc_specifyParameters(thisIntegrator%handle, p), thisIntegrator%handle, status, "specifyParameters")
    if (present(intensityMus) .and. .not. stateIsFailure(status)) thisIntegrator%nDirections = size(intensityMus)
  end subroutine

  ! ---- computeRadiativeTransfer (monteCarloRadiativeTransfer.f95:262-398)
  ! The photonStream of the replacement module monteCarloIllumination carries a descriptor (kind + parameters); a stream
  ! whose public arrays were filled by hand is passed as I3RC_SRC_ARRAYS (descriptorOf).  The seed vector that created
  ! randomNumbers keys the per-photon Philox streams: the replacement module RandomNumbers keeps it in the type
  ! (randomNumbers%seed, %nSeed), because the twister's state holds no recoverable trace of it.
  subroutine computeRadiativeTransfer(thisIntegrator, randomNumbers, incomingPhotons, status)
    type(integrator), intent(inout) :: thisIntegrator
    type(randomNumberSequence), intent(inout) :: randomNumbers
    type(photonStream), intent(inout) :: incomingPhotons
    type(ErrorMessage), intent(inout) :: status
    type(i3rc_photon_source) :: src
    integer(c_int32_t) :: seed(2)
    src  = descriptorOf(incomingPhotons)
    seed = randomNumbers%seed
    call toStatus(i3rc_computeRadiativeTransfer(thisIntegrator%handle, src, seed, min(max(randomNumbers%nSeed, 1), 2)), &
                  thisIntegrator%handle, status, "computeRadiativeTransfer")
    if (.not. stateIsFailure(status)) then
      call setStateToCompleteSuccess(status, "computeRadiativeTransfer: finished with photons")
      incomingPhotons%currentPhoton = int(src%numberOfPhotons) + 1      ! the stream is used up (MCRT:472, 700)
    end if
  end subroutine

  ! ---- reportResults (monteCarloRadiativeTransfer.f95:711-826): absent optionals -> NULL; sizes checked like :746-791
  subroutine reportResults(thisIntegrator, meanFluxUp, meanFluxDown, meanFluxAbsorbed, fluxUp, fluxDown, fluxAbsorbed, &
                           absorbedProfile, volumeAbsorption, meanIntensity, intensity, status)
    type(integrator), intent(in) :: thisIntegrator
    real, optional, target, intent(out) :: meanFluxUp, meanFluxDown, meanFluxAbsorbed
    real, dimension(:, :), optional, target, contiguous, intent(out) :: fluxUp, fluxDown, fluxAbsorbed
    real, dimension(:), optional, target, contiguous, intent(out) :: absorbedProfile, meanIntensity
    real, dimension(:, :, :), optional, target, contiguous, intent(out) :: volumeAbsorption, intensity
    type(ErrorMessage), intent(inout) :: status
    type(c_ptr) :: p(10)
    integer :: nX, nY, nZ, nD
    nX = thisIntegrator%nX; nY = thisIntegrator%nY; nZ = thisIntegrator%nZ; nD = thisIntegrator%nDirections
    p(:) = c_null_ptr
    if (present(meanFluxUp)) p(1) = c_loc(meanFluxUp)
    if (present(meanFluxDown)) p(2) = c_loc(meanFluxDown)
    if (present(meanFluxAbsorbed)) p(3) = c_loc(meanFluxAbsorbed)
    if (present(fluxUp)) then
      if (any(shape(fluxUp) /= (/ nX, nY /))) call setStateToFailure(status, "reportResults: fluxUp array is the wrong size")
      p(4) = c_loc(fluxUp)
    end if
    if (present(fluxDown)) then
      if (any(shape(fluxDown) /= (/ nX, nY /))) call setStateToFailure(status, "reportResults: fluxDown array is the wrong size")
      p(5) = c_loc(fluxDown)
    end if
    if (present(fluxAbsorbed)) then
      if (any(shape(fluxAbsorbed) /= (/ nX, nY /))) &
        call setStateToFailure(status, "reportResults: fluxAbsorbed array is the wrong size")
      p(6) = c_loc(fluxAbsorbed)
    end if
    if (present(absorbedProfile)) then
      if (size(absorbedProfile) /= nZ) call setStateToFailure(status, "reportResults: absorbedProfile array is the wrong size")
      p(7) = c_loc(absorbedProfile)
    end if
    if (present(volumeAbsorption)) then
      if (any(shape(volumeAbsorption) /= (/ nX, nY, nZ /))) &
        call setStateToFailure(status, "reportResults: volumeAbsorption array is the wrong size")
      p(8) = c_loc(volumeAbsorption)
    end if
    if (present(meanIntensity)) then
      if (nD == 0) call setStateToFailure(status, "reportResults: intensity information not available")
      if (size(meanIntensity) /= nD) call setStateToFailure(status, "reportResults: requesting mean intensity in the wrong number of directions.")
      p(9) = c_loc(meanIntensity)
    end if
    if (present(intensity)) then
      if (nD == 0) call setStateToFailure(status, "reportResults: intensity information not available")
      if (any(shape(intensity) /= (/ nX, nY, nD /))) call setStateToFailure(status, "reportResults: intensity array has wrong dimensions.")
      p(10) = c_loc(intensity)
    end if
    if (stateIsFailure(status)) return
    call toStatus(i3rc_reportResults(thisIntegrator%handle, p(1), p(2), p(3), p(4), p(5), p(6), p(7), p(8), p(9), p(10)), &
                  thisIntegrator%handle, status, "reportResults")
  end subroutine

  function copy_Integrator(original) result(copy)
    type(integrator), intent(in) :: original; type(integrator) :: copy
    integer(c_int) :: rc
    copy = original; rc = i3rc_copy_Integrator(original%handle, copy%handle)
  end function
  function isReady_Integrator(thisIntegrator)
    type(integrator), intent(in) :: thisIntegrator; logical :: isReady_Integrator
    isReady_Integrator = c_associated(thisIntegrator%handle)
    if (isReady_Integrator) isReady_Integrator = i3rc_isReady_Integrator(thisIntegrator%handle) /= 0
  end function
  subroutine finalize_Integrator(thisIntegrator)
    type(integrator), intent(inout) :: thisIntegrator
    if (c_associated(thisIntegrator%handle)) call i3rc_finalize_Integrator(thisIntegrator%handle)
    thisIntegrator%handle = c_null_ptr
  end subroutine
end module monteCarloRadiativeTransfer
