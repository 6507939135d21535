! Replacement for Code/monteCarloIllumination.f95: same module name, same public list (monteCarloIllumination.f95:54-55),
! same six constructors behind the generic new_PhotonStream (monteCarloIllumination.f95:46-50) with the same argument
! lists, checks and status texts (monteCarloIllumination.f95:78-83, 122-125, 163-164, 200-208, 251-268, 353-376).
!
! What changes: a constructor no longer draws numberOfPhotons positions on the host.  It records WHAT was asked for in
! photons%descriptor (the C struct i3rc_photon_source of include/i3rc_b200.h); the CUDA integrator draws every photon on
! the device from that photon's own Philox stream (csrc/transport.cuh, init_photon_state).  The five public array
! components of type(photonStream) stay (monteCarloIllumination.f95:34-41): a caller that allocates and fills them by
! hand gets descriptor%kind = I3RC_SRC_ARRAYS and computeRadiativeTransfer uploads the arrays (materialize_PhotonStream
! does that for a descriptor stream, on the host, for callers that want to walk the photons with getNextPhoton).
! randomNumbers is accepted and left untouched by the descriptor constructors: the sequence that is later handed to
! computeRadiativeTransfer keys the device streams (module RandomNumbers of this directory keeps its seed).
!
! NOT COMPILED in the build environment (no Fortran compiler exists there).  tests/ctest/shim_replay.c makes the same
! C calls in the same order and is compiled and run by the test-suite.
module monteCarloIllumination
  use, intrinsic :: iso_c_binding
  use ErrorMessages, only: ErrorMessage, stateIsFailure, setStateToFailure, setStateToWarning, setStateToSuccess
  use RandomNumbers, only: randomNumberSequence, getRandomReal
  implicit none
  private

  ! include/i3rc_b200.h: enum i3rc_source_kind and struct i3rc_photon_source
  integer(c_int32_t), parameter, public :: I3RC_SRC_DIRECTIONAL = 1, I3RC_SRC_RANDOM_AZIMUTH = 2, I3RC_SRC_FLUX = 3,   &
                                           I3RC_SRC_SPOTLIGHT = 4, I3RC_SRC_INTERNAL_FLUX = 5,                         &
                                           I3RC_SRC_INTERNAL_INTENSITY = 6, I3RC_SRC_ARRAYS = 7
  type, bind(c), public :: i3rc_photon_source
    integer(c_int32_t) :: kind = 0, reserved = 0
    integer(c_int64_t) :: numberOfPhotons = 0
    real(c_float)      :: solarMu = 0., solarAzimuth = 0., x = 0., y = 0., z = 0., detectorMu = 0., detectorPhi = 0.
    integer(c_int32_t) :: detectorPointsUp = 0, has_deltaX = 0, has_deltaY = 0
    real(c_float)      :: deltaX = 0., deltaY = 0.
    type(c_ptr)        :: xPosition = c_null_ptr, yPosition = c_null_ptr, zPosition = c_null_ptr, &
                          initialMu = c_null_ptr, initialPhi = c_null_ptr
  end type i3rc_photon_source

  type photonStream
    integer                     :: currentPhoton = 0
    real, dimension(:), pointer :: xPosition => null()
    real, dimension(:), pointer :: yPosition => null()
    real, dimension(:), pointer :: zPosition => null()
    real, dimension(:), pointer :: initialMu  => null()
    real, dimension(:), pointer :: initialPhi => null()
    type(i3rc_photon_source)    :: descriptor          ! what the constructor was asked for (kind 0: none)
  end type photonStream

  interface new_PhotonStream
    module procedure newPhotonStream_Directional, newPhotonStream_RandomAzimuth, &
                     newPhotonStream_Flux, newPhotonStream_Spotlight,            &
                     newPhotonStream_Internal_Flux, newPhotonStream_Internal_Intensity
  end interface new_PhotonStream

  public :: photonStream
  public :: new_PhotonStream, finalize_PhotonStream, morePhotonsExist, getNextPhoton
  public :: materialize_PhotonStream, descriptorOf     ! additions (used by module monteCarloRadiativeTransfer)
contains
  ! ---- the six constructors ------------------------------------------------------------------------------------------
  function newPhotonStream_Directional(solarMu, solarAzimuth, numberOfPhotons, randomNumbers, status) result(photons)
    real,                       intent(in   ) :: solarMu, solarAzimuth
    integer                                   :: numberOfPhotons
    type(randomNumberSequence), intent(inout) :: randomNumbers
    type(ErrorMessage),         intent(inout) :: status
    type(photonStream)                        :: photons
    if(numberOfPhotons <= 0) call setStateToFailure(status, "setIllumination: must ask for non-negative number of photons.")
    if(solarAzimuth < 0. .or. solarAzimuth > 360.) call setStateToFailure(status, "setIllumination: solarAzimuth out of bounds")
    if(abs(solarMu) > 1. .or. abs(solarMu) <= tiny(solarMu)) call setStateToFailure(status, "setIllumination: solarMu out of bounds")
    if(stateIsFailure(status)) return
    photons%descriptor%kind = I3RC_SRC_DIRECTIONAL
    photons%descriptor%numberOfPhotons = numberOfPhotons
    photons%descriptor%solarMu = solarMu;  photons%descriptor%solarAzimuth = solarAzimuth
    photons%currentPhoton = 1
    call setStateToSuccess(status)
  end function newPhotonStream_Directional

  function newPhotonStream_RandomAzimuth(solarMu, numberOfPhotons, randomNumbers, status) result(photons)
    real,                       intent(in   ) :: solarMu
    integer                                   :: numberOfPhotons
    type(randomNumberSequence), intent(inout) :: randomNumbers
    type(ErrorMessage),         intent(inout) :: status
    type(photonStream)                        :: photons
    if(numberOfPhotons <= 0) call setStateToFailure(status, "setIllumination: must ask for non-negative number of photons.")
    if(abs(solarMu) > 1. .or. abs(solarMu) <= tiny(solarMu)) call setStateToFailure(status, "setIllumination: solarMu out of bounds")
    if(stateIsFailure(status)) return
    photons%descriptor%kind = I3RC_SRC_RANDOM_AZIMUTH
    photons%descriptor%numberOfPhotons = numberOfPhotons
    photons%descriptor%solarMu = solarMu
    photons%currentPhoton = 1
    call setStateToSuccess(status)
  end function newPhotonStream_RandomAzimuth

  function newPhotonStream_Flux(numberOfPhotons, randomNumbers, status) result(photons)
    integer                                   :: numberOfPhotons
    type(randomNumberSequence), intent(inout) :: randomNumbers
    type(ErrorMessage),         intent(inout) :: status
    type(photonStream)                        :: photons
    if(numberOfPhotons <= 0) call setStateToFailure(status, "setIllumination: must ask for non-negative number of photons.")
    if(stateIsFailure(status)) return
    photons%descriptor%kind = I3RC_SRC_FLUX
    photons%descriptor%numberOfPhotons = numberOfPhotons
    photons%currentPhoton = 1
    call setStateToSuccess(status)
  end function newPhotonStream_Flux

  function newPhotonStream_Spotlight(solarMu, solarAzimuth, solarX, solarY, numberOfPhotons, randomNumbers, status) &
           result(photons)
    real,                       intent(in   ) :: solarMu, solarAzimuth, solarX, solarY
    integer                                   :: numberOfPhotons
    type(randomNumberSequence), optional, intent(inout) :: randomNumbers
    type(ErrorMessage),         intent(inout) :: status
    type(photonStream)                        :: photons
    if(numberOfPhotons <= 0) call setStateToFailure(status, "setIllumination: must ask for non-negative number of photons.")
    if(solarAzimuth < 0. .or. solarAzimuth > 360.) call setStateToFailure(status, "setIllumination: solarAzimuth out of bounds")
    if(abs(solarMu) > 1. .or. abs(solarMu) <= tiny(solarMu)) call setStateToFailure(status, "setIllumination: solarMu out of bounds")
    if(solarX > 1. .or. solarX <= 0. .or. solarY > 1. .or. solarY <= 0. ) &
      call setStateToFailure(status, "setIllumination: x and y positions must be between 0 and 1")
    if(stateIsFailure(status)) return
    photons%descriptor%kind = I3RC_SRC_SPOTLIGHT
    photons%descriptor%numberOfPhotons = numberOfPhotons
    photons%descriptor%solarMu = solarMu;  photons%descriptor%solarAzimuth = solarAzimuth
    photons%descriptor%x = solarX;         photons%descriptor%y = solarY
    photons%currentPhoton = 1
    call setStateToSuccess(status)
  end function newPhotonStream_Spotlight

  function newPhotonStream_Internal_Flux(detectorX, detectorY, detectorZ, detectorPointsUp, deltaX, deltaY, &
                                         numberOfPhotons, randomNumbers, status) result(photons)
    real,                       intent(in)    :: detectorX, detectorY, detectorZ
    logical,                    intent(in)    :: detectorPointsUp
    real,             optional, intent(in)    :: deltaX, deltaY
    integer                                   :: numberOfPhotons
    type(randomNumberSequence), optional, intent(inout) :: randomNumbers
    type(ErrorMessage),         intent(inout) :: status
    type(photonStream)                        :: photons
    call checkDetector(detectorX, detectorY, detectorZ, deltaX, deltaY, numberOfPhotons, status)
    if(      detectorPointsUp .and. abs(detectorZ - 1.) < 2. * spacing(1.)) &
      call setStateToWarning(status, "setIllumination: Detector is at top of domain pointed up")
    if(.not. detectorPointsUp .and. detectorZ           < 2. * tiny(0.)   ) &
      call setStateToWarning(status, "setIllumination: Detector is at bottom of domain pointed down")
    if(stateIsFailure(status)) return
    photons%descriptor%kind = I3RC_SRC_INTERNAL_FLUX
    photons%descriptor%numberOfPhotons = numberOfPhotons
    photons%descriptor%x = detectorX;  photons%descriptor%y = detectorY;  photons%descriptor%z = detectorZ
    photons%descriptor%detectorPointsUp = merge(1, 0, detectorPointsUp)
    call setWidths(photons%descriptor, deltaX, deltaY)
    photons%currentPhoton = 1
  end function newPhotonStream_Internal_Flux

  function newPhotonStream_Internal_Intensity(detectorX, detectorY, detectorZ, detectorMu, detectorPhi,  &
                                              deltaX, deltaY, deltaTheta, numberOfPhotons, randomNumbers, status) &
           result(photons)
    real,               intent(in)    :: detectorX, detectorY, detectorZ
    real,               intent(in)    :: detectorMu, detectorPhi     ! phi in degrees
    real,     optional, intent(in)    :: deltaX, deltaY, deltaTheta  ! (deltaTheta is accepted and unused, like the reference)
    integer                           :: numberOfPhotons
    type(randomNumberSequence), optional, intent(inout) :: randomNumbers
    type(ErrorMessage), intent(inout) :: status
    type(photonStream)                :: photons
    call checkDetector(detectorX, detectorY, detectorZ, deltaX, deltaY, numberOfPhotons, status)
    if(detectorPhi < 0. .or. detectorPhi > 360.) call setStateToFailure(status, "setIllumination: detectorPhi out of bounds")
    if(abs(detectorMu) > 1. .or. abs(detectorMu) <= tiny(detectorMu)) &
      call setStateToFailure(status, "setIllumination: detectorMu out of bounds")
    if(stateIsFailure(status)) return
    photons%descriptor%kind = I3RC_SRC_INTERNAL_INTENSITY
    photons%descriptor%numberOfPhotons = numberOfPhotons
    photons%descriptor%x = detectorX;  photons%descriptor%y = detectorY;  photons%descriptor%z = detectorZ
    photons%descriptor%detectorMu = detectorMu;  photons%descriptor%detectorPhi = detectorPhi
    call setWidths(photons%descriptor, deltaX, deltaY)
    photons%currentPhoton = 1
  end function newPhotonStream_Internal_Intensity

  ! checks shared by the two internal sources (monteCarloIllumination.f95:251-264, 353-366)
  subroutine checkDetector(detectorX, detectorY, detectorZ, deltaX, deltaY, numberOfPhotons, status)
    real,               intent(in)    :: detectorX, detectorY, detectorZ
    real,     optional, intent(in)    :: deltaX, deltaY
    integer,            intent(in)    :: numberOfPhotons
    type(ErrorMessage), intent(inout) :: status
    if(numberOfPhotons <= 0) call setStateToFailure(status, "setIllumination: must ask for non-negative number of photons.")
    if(detectorX > 1. .or. detectorX <= 0. .or. detectorY > 1. .or. detectorY <= 0. .or. &
       detectorZ > 1. .or. detectorZ <= 0. ) &
      call setStateToFailure(status, "setIllumination: x, y, z positions must be between 0 and 1")
    if(present(deltaX)) then
      if(detectorX + deltaX/2. > 1. .or. detectorX - deltaX/2 <= 0.) &
        call setStateToFailure(status, "setIllumination: max, min positions must be between 0 and 1")
    end if
    if(present(deltaY)) then
      if(detectorY + deltaY/2. > 1. .or. detectorY - deltaY/2 <= 0.) &
        call setStateToFailure(status, "setIllumination: max, min positions must be between 0 and 1")
    end if
  end subroutine checkDetector
  subroutine setWidths(d, deltaX, deltaY)
    type(i3rc_photon_source), intent(inout) :: d
    real, optional,           intent(in)    :: deltaX, deltaY
    if(present(deltaX)) then;  d%has_deltaX = 1;  d%deltaX = deltaX;  end if
    if(present(deltaY)) then;  d%has_deltaY = 1;  d%deltaY = deltaY;  end if
  end subroutine setWidths

  ! ---- what computeRadiativeTransfer hands to the C ABI ---------------------------------------------------------------
  ! A stream made by a constructor: its descriptor.  A stream whose public arrays were allocated and filled by the caller
  ! (or by materialize_PhotonStream): kind = I3RC_SRC_ARRAYS with the addresses of the five arrays.
  function descriptorOf(photons) result(d)
    type(photonStream), target, intent(in) :: photons
    type(i3rc_photon_source) :: d
    d = photons%descriptor
    if(associated(photons%xPosition)) then
      d%kind = I3RC_SRC_ARRAYS
      d%numberOfPhotons = size(photons%xPosition)
      d%xPosition  = c_loc(photons%xPosition(1));  d%yPosition  = c_loc(photons%yPosition(1))
      d%zPosition  = c_loc(photons%zPosition(1));  d%initialMu  = c_loc(photons%initialMu(1))
      d%initialPhi = c_loc(photons%initialPhi(1))
    end if
  end function descriptorOf

  ! Draw the photons of a descriptor stream on the HOST into the public arrays, with the reference's recipes and its
  ! order of draws (monteCarloIllumination.f95:91-99, 132-140, 172-181, 213-219): for callers that walk a stream with
  ! getNextPhoton.  The two internal (backward Monte Carlo) sources are drawn on the device only.
  subroutine materialize_PhotonStream(photons, randomNumbers, status)
    type(photonStream),         intent(inout) :: photons
    type(randomNumberSequence), intent(inout) :: randomNumbers
    type(ErrorMessage),         intent(inout) :: status
    integer :: i, n
    real    :: twoPi
    twoPi = 2. * acos(-1.)
    n = int(photons%descriptor%numberOfPhotons)
    if(photons%descriptor%kind < I3RC_SRC_DIRECTIONAL .or. photons%descriptor%kind > I3RC_SRC_SPOTLIGHT .or. n <= 0) then
      call setStateToFailure(status, "materialize_PhotonStream: nothing to draw on the host for this stream.")
      return
    end if
    call finalize_arrays(photons)
    allocate(photons%xPosition(n), photons%yPosition(n), photons%zPosition(n), photons%initialMu(n), photons%initialPhi(n))
    photons%zPosition(:)  = 1. - spacing(1.)
    photons%initialMu(:)  = -abs(photons%descriptor%solarMu)
    photons%initialPhi(:) = photons%descriptor%solarAzimuth * acos(-1.) / 180.
    select case(photons%descriptor%kind)
      case(I3RC_SRC_DIRECTIONAL)
        do i = 1, n
          photons%xPosition(i) = getRandomReal(randomNumbers);  photons%yPosition(i) = getRandomReal(randomNumbers)
        end do
      case(I3RC_SRC_RANDOM_AZIMUTH)
        do i = 1, n
          photons%xPosition(i) = getRandomReal(randomNumbers);  photons%yPosition(i) = getRandomReal(randomNumbers)
          photons%initialPhi(i) = getRandomReal(randomNumbers) * twoPi
        end do
      case(I3RC_SRC_FLUX)
        do i = 1, n
          photons%xPosition(i) = getRandomReal(randomNumbers);  photons%yPosition(i) = getRandomReal(randomNumbers)
          photons%initialMu(i)  = -sqrt(getRandomReal(randomNumbers))
          photons%initialPhi(i) = getRandomReal(randomNumbers) * twoPi
        end do
      case(I3RC_SRC_SPOTLIGHT)
        photons%xPosition(:) = photons%descriptor%x;  photons%yPosition(:) = photons%descriptor%y
    end select
    photons%currentPhoton = 1
    call setStateToSuccess(status)
  end subroutine materialize_PhotonStream

  ! ---- walking a stream (monteCarloIllumination.f95:428-459) ----------------------------------------------------------
  function morePhotonsExist(photons)
    type(photonStream), intent(inout) :: photons
    logical                           :: morePhotonsExist
    if(associated(photons%xPosition)) then
      morePhotonsExist = photons%currentPhoton > 0 .and. photons%currentPhoton <= size(photons%xPosition)
    else
      morePhotonsExist = photons%currentPhoton > 0 .and. photons%currentPhoton <= photons%descriptor%numberOfPhotons
    end if
  end function morePhotonsExist

  subroutine getNextPhoton(photons, xPosition, yPosition, zPosition, solarMu, solarAzimuth, status)
    type(photonStream), intent(inout) :: photons
    real,               intent(  out) :: xPosition, yPosition, zPosition, solarMu, solarAzimuth
    type(ErrorMessage), intent(inout) :: status
    if(photons%currentPhoton < 1) call setStateToFailure(status, "getNextPhoton: photons have not been initialized.")
    if(.not. stateIsFailure(status) .and. .not. associated(photons%xPosition)) &
      call setStateToFailure(status, "getNextPhoton: the photons of this stream are drawn on the device" // &
                                     " (call materialize_PhotonStream to draw them on the host).")
    if(.not. stateIsFailure(status)) then
      if(photons%currentPhoton > size(photons%xPosition)) call setStateToFailure(status, "getNextPhoton: Ran out of photons")
    end if
    if(.not. stateIsFailure(status)) then
      xPosition    = photons%xPosition(photons%currentPhoton)
      yPosition    = photons%yPosition(photons%currentPhoton)
      zPosition    = photons%zPosition(photons%currentPhoton)
      solarMu      = photons%initialMu(photons%currentPhoton)
      solarAzimuth = photons%initialPhi(photons%currentPhoton)
      photons%currentPhoton = photons%currentPhoton + 1
    end if
  end subroutine getNextPhoton

  ! ---- finalization (monteCarloIllumination.f95:463-474) ---------------------------------------------------------------
  subroutine finalize_arrays(photons)
    type(photonStream), intent(inout) :: photons
    if(associated(photons%xPosition))  deallocate(photons%xPosition)
    if(associated(photons%yPosition))  deallocate(photons%yPosition)
    if(associated(photons%zPosition))  deallocate(photons%zPosition)
    if(associated(photons%initialMu))  deallocate(photons%initialMu)
    if(associated(photons%initialPhi)) deallocate(photons%initialPhi)
  end subroutine finalize_arrays
  subroutine finalize_PhotonStream(photons)
    type(photonStream), intent(inout) :: photons
    type(i3rc_photon_source) :: none
    call finalize_arrays(photons)
    photons%descriptor = none
    photons%currentPhoton = 0
  end subroutine finalize_PhotonStream
end module monteCarloIllumination
