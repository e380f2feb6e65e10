// Host mirror of main/src/utilities/timers.h:9-51: the two stopwatch classes scene code prints with.
// hostTimer is std::chrono based like the reference's; cudaTimer has no device work to bracket on the
// host side of the ABI (kernel time is reported by rtb_get_counters), so it measures wall time too.
#pragma once
#include <chrono>

class hostTimer {
	std::chrono::steady_clock::time_point t0{}, t1{};

public:
	void start() { t0 = std::chrono::steady_clock::now(); }
	void end() { t1 = std::chrono::steady_clock::now(); }
	float elapsedms() const { return std::chrono::duration<float, std::milli>(t1 - t0).count(); }
};

class cudaTimer : public hostTimer {};
