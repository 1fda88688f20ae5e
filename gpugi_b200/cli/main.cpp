// gpugi_b200_cli -- runs a gpugi command script (scenes/*.txt syntax) headless: the role of gpugi/main.cpp:15-62 without the window.
//   gpugi_b200_cli <script.txt> [cuda device]
// Links the host layer through its flat C view (include/gpugi_b200_host.h): Application(argc, argv) + Application::Run().
#include "../../include/gpugi_b200_host.h"

#include <cstdlib>
#include <iostream>

int main(int argc, char** argv)
{
	if (argc < 2)
	{
		std::cerr << "usage: " << argv[0] << " <script.txt> [cuda device]\n";
		return 2;
	}
	GihApplication* app = gih_app_create(argv[1], argc > 2 ? std::atoi(argv[2]) : 0);
	if (!app)
	{
		std::cerr << "Unhandled exception: " << gih_last_error() << std::endl;
		return 1;
	}
	gih_app_run(app, 0);
	gih_app_destroy(app); // writes a last screenshot unless the script ended with "exit" (application.cpp:196-198)
	return 0;
}
