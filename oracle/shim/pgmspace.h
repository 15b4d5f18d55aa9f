/* host stand-in for the Arduino header the reference's libopus copy includes (celt/cwrs.c:38,
 * silk/VAD.c:31, silk/sigm_Q15.c:31): PROGMEM places tables in flash on the ESP32 and is empty on a host */
#ifndef ORACLE_SHIM_PGMSPACE_H
#define ORACLE_SHIM_PGMSPACE_H
#define PROGMEM
#endif
